"""Programmatic dependent launch (csrc/common.cuh launch_k / pdl_wait) must not change a single bit: a kernel launched with the attribute
starts early but blocks in griddepcontrol.wait until its predecessor has completed.  The switch is read once per process, so each
setting runs in its own interpreter: two training steps of the miniature network (both dtypes), SHA-256 of the parameters and of the
last prediction, for RESNET_B200_PDL = 0 (every launch serialized), 2 (default: convolution kernels), 7 (every kernel class)."""
import hashlib  # noqa: F401  (used by the child)
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r"""
import hashlib, sys
import numpy as np
sys.path.insert(0, %r)
from oracle import golden_cases as G
from oracle import oracle as O
from resnet_b200 import api
cfg = dict(G.MINI, batch=8, input_dim=64)
shapes = O.param_shapes(cfg["input_dim"], cfg["n_blocks"], cfg["reductions"], output=cfg["output"])
W = G.mini_weights(shapes)
img, lab = G.mini_batch(cfg)
for dtype in ("tf32", "bf16"):
    t = api.Trainer(input_dim=cfg["input_dim"], n_blocks=cfg["n_blocks"], reductions=cfg["reductions"], batch=cfg["batch"],
                    output=cfg["output"], lr=cfg["lr"], dtype=dtype)
    t.set_params(W)
    h = hashlib.sha256()
    for _ in range(2):
        t.set_batch(img, lab)
        pred = t.forward()
        t.backward()
        t.update()
    t.sync()
    h.update(np.ascontiguousarray(pred).tobytes())
    for p in t.get_params(0):
        h.update(np.ascontiguousarray(p).tobytes())
    assert np.isfinite(pred).all()
    print("HASH", dtype, h.hexdigest())
    t.close()
""" % ROOT


def run(mask):
    env = dict(os.environ, RESNET_B200_PDL=str(mask))
    r = subprocess.run([sys.executable, "-c", CHILD], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=240)
    assert r.returncode == 0, r.stderr[-2000:]
    return [line for line in r.stdout.splitlines() if line.startswith("HASH")]


def test_pdl_modes_bit_identical():
    ref = run(0)
    assert len(ref) == 2
    for mask in (2, 7):
        assert run(mask) == ref, "RESNET_B200_PDL=%d changed the results" % mask
