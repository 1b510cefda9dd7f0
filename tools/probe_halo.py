"""Haloed-patch 3x3 kernel (igemm_halo_kernel) against the per-tap kernel.
    python tools/probe_halo.py check   # small problems: halo (descriptor base offset off / on) vs per-tap kernel vs the host oracle
    python tools/probe_halo.py time    # full-size layers, one launch per printed line, for an ncu launch list"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from resnet_b200 import api  # noqa: E402
from oracle import oracle as O  # noqa: E402  (checker only)


def rel_max(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def run(x, w, dy, dtype, env, base=None):
    keys = ("RESNET_B200_HALO", "RESNET_B200_HALO_BASEOFF", "RESNET_B200_RESIDENT_B", "RESNET_B200_OP_STATS")
    for k in keys:
        os.environ.pop(k, None)
    os.environ.update(env)
    try:
        y = api.conv_forward(x, w, 1, impl=0, dtype=dtype)
        din, _ = api.conv_backward(x, w, dy, 1, din_base=base, impl=0, dtype=dtype)
    finally:
        for k in keys:
            os.environ.pop(k, None)
    return y, din


def check():
    rng = np.random.default_rng(0)
    shapes = [(56, 64, 64, 3), (28, 128, 128, 3), (20, 64, 64, 3), (12, 128, 64, 2), (16, 64, 384, 2), (28, 64, 256, 2), (8, 64, 64, 5)]
    for (S, cin, cout, N) in shapes:
        for dtype in ("f32", "bf16"):
            R = api.bf16_round if dtype == "bf16" else (lambda a: a)
            x = R(rng.standard_normal((N, S, S, cin)).astype(np.float32))
            w = R((rng.standard_normal((cout, cin, 3, 3)) * 0.1).astype(np.float32))
            dy = R(rng.standard_normal((N, S, S, cout)).astype(np.float32))
            base = R(rng.standard_normal((N, S, S, cin)).astype(np.float32))
            y0, d0 = run(x, w, dy, dtype, dict(RESNET_B200_HALO="0"))
            yo, do = O.conv_fwd(x, w, 1), O.conv_dgrad(w, dy, S, 1)
            line = "3x3 %d->%d @%d N=%d %-4s per-tap vs oracle y %.1e dx %.1e |" % (cin, cout, S, N, dtype, rel_max(y0, yo), rel_max(d0, do))
            for name, env in (("halo", dict(RESNET_B200_HALO="2")), ("halo+baseoff", dict(RESNET_B200_HALO="2", RESNET_B200_HALO_BASEOFF="1")),
                              ("halo resB=2", dict(RESNET_B200_HALO="2", RESNET_B200_RESIDENT_B="2")),
                              ("halo resB=0", dict(RESNET_B200_HALO="2", RESNET_B200_RESIDENT_B="0"))):
                y1, d1 = run(x, w, dy, dtype, env)
                line += " %s: y %.1e dx %.1e |" % (name, rel_max(y1, y0), rel_max(d1, d0))
            # dgrad accumulating into an existing gradient (the residual join, TMA reduce-add)
            _, da0 = run(x, w, dy, dtype, dict(RESNET_B200_HALO="0"), base=base)
            _, da1 = run(x, w, dy, dtype, dict(RESNET_B200_HALO="2"), base=base)
            line += " accumulate dx %.1e" % rel_max(da1, da0)
            print(line, flush=True)


def timing():
    rng = np.random.default_rng(0)
    for (S, cin, cout, N) in [(56, 64, 64, 256), (28, 128, 128, 256)]:
        x = rng.standard_normal((N, S, S, cin), dtype=np.float32)
        w = rng.standard_normal((cout, cin, 3, 3), dtype=np.float32) * 0.05
        for dtype in ("bf16", "f32"):
            for env in (dict(RESNET_B200_HALO="0"), dict(RESNET_B200_HALO="1"), dict(RESNET_B200_HALO="1", RESNET_B200_OP_STATS="1"),
                        dict(RESNET_B200_HALO="1", RESNET_B200_RESIDENT_B="0"), dict(RESNET_B200_HALO="1", RESNET_B200_EPI_GROUPS="1"),
                        dict(RESNET_B200_HALO="1", RESNET_B200_EPI_GROUPS="2")):
                for k in ("RESNET_B200_HALO", "RESNET_B200_RESIDENT_B", "RESNET_B200_OP_STATS", "RESNET_B200_EPI_GROUPS"):
                    os.environ.pop(k, None)
                os.environ.update(env)
                y = api.conv_forward(x, w, 1, impl=0, dtype=dtype)
                print("fprop 3x3/1 %d->%d @%d %s %s mean|y|=%.4f" % (cin, cout, S, dtype, " ".join("%s=%s" % (k[12:], v) for k, v in env.items()),
                                                                   float(np.abs(y).mean())), flush=True)


if __name__ == "__main__":
    (check if sys.argv[1:] == ["check"] else timing)()
