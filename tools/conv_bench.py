"""Per-layer timing of the tcgen05 convolution kernels with CUDA events (no ncu needed): every distinct convolution shape of the
reference's ResNet-50 x {fprop + fused statistics, dgrad, wgrad + reduce}, at the bench batch size, through
resnet_b200_conv_bench (plan built once, launches back to back on one stream).

    python tools/conv_bench.py [--dtype bf16|f32] [--batch 256] [--iters 10] [--only 3x3] [--env K=V ...] [--variants "A=1;B=2,C=3"]

--variants: semicolon-separated environment settings (comma-separated K=V inside one variant) to compare against the default plan, one
column each.  Prints microseconds, algorithmic TFLOP/s (2*N*Ho*Wo*Cout*Cin*k^2) and the fraction of the measured tensor peak
(MEASURED_PEAKS.json bf16_tflops burst for a kernel timed alone; half of it for TF32).
"""
import argparse
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from resnet_b200 import api  # noqa: E402


def r50_shapes():
    """(S, k, cin, cout, stride, count, has_dgrad): distinct convolutions of the reference's ResNet-50 variant (stride on the 3x3,
    3x3/2 projections; resnet.cu:3251-3258) with how many times each occurs per step"""
    shapes = {}

    def add(S, k, cin, cout, stride, dgrad=True):
        key = (S, k, cin, cout, stride, dgrad)
        shapes[key] = shapes.get(key, 0) + 1

    add(224, 7, 3, 64, 2, False)
    incoming, spatial, reduced, expanded = 64, 56, 64, 256
    for i in range(16):
        stride = 1
        if i in (3, 7, 13):
            stride, reduced, expanded = 2, reduced * 2, expanded * 2
        add(spatial, 1, incoming, reduced, 1)
        add(spatial, 3, reduced, reduced, stride)
        add(spatial // stride, 1, reduced, expanded, 1)
        if incoming != expanded:
            add(spatial, 3 if stride == 2 else 1, incoming, expanded, stride)
        spatial //= stride
        incoming = expanded
    return [(k + (v,)) for k, v in shapes.items()]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dtype", default="bf16")
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--only", default="")
    ap.add_argument("--shape", default="", help="k,stride,cin,cout,S: only this layer (shell-safe form of --only)")
    ap.add_argument("--passes", default="0,1,3")
    ap.add_argument("--variants", default="")
    ap.add_argument("--desc", action="store_true")
    args = ap.parse_args()
    L = api.L()
    bf = args.dtype == "bf16"
    L.resnet_b200_set_op_dtype(int(bf))
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = peaks.get("bf16_tflops", 1590.0) * (1.0 if bf else 0.5)
    variants = [dict()] + [dict(kv.split("=") for kv in v.split(",") if kv) for v in args.variants.split(";") if v]
    names = {0: "fprop+st", 1: "dgrad", 2: "dgrad+=", 3: "wgrad"}
    passes = [int(p) for p in args.passes.split(",")]
    total = [0.0] * len(variants)
    total_fl = 0.0
    print("# %s batch %d, peak %.0f TFLOP/s (%s); columns: us TFLOP/s frac per variant %s" % (args.dtype, args.batch, peak, "measured" if peaks else "fallback",
                                                                                             [" ".join("%s=%s" % kv for kv in v.items()) or "default" for v in variants]))
    desc = C.create_string_buffer(512)
    for (S, k, cin, cout, stride, dgrad, count) in r50_shapes():
        tag = "%dx%d/%d %d->%d @%d" % (k, k, stride, cin, cout, S)
        if args.only and args.only not in tag:
            continue
        if args.shape and [int(v) for v in args.shape.split(",")] != [k, stride, cin, cout, S]:
            continue
        So = S // stride
        fl = 2.0 * args.batch * So * So * cout * cin * k * k
        for ps in passes:
            if ps in (1, 2) and not dgrad:
                continue
            line = "%-8s %-26s x%-2d" % (names[ps], tag, count)
            for vi, v in enumerate(variants):
                saved = {kk: os.environ.get(kk) for kk in v}
                os.environ.update(v)
                ms = L.resnet_b200_conv_bench(S, k, cin, cout, stride, args.batch, ps, 1, args.warmup, args.iters, desc, 512)
                for kk, old in saved.items():
                    if old is None:
                        os.environ.pop(kk, None)
                    else:
                        os.environ[kk] = old
                err = L.resnet_b200_last_error().decode()
                if ms < 0 or err:
                    line += " |   FAILED %s" % err[:60]
                    L.resnet_b200_clear_error()
                    continue
                tf = fl / (ms * 1e-3) / 1e12
                line += " | %8.1f %7.1f %5.2f" % (ms * 1e3, tf, tf / peak)
                total[vi] += ms * count
                if args.desc:
                    line += "  [%s]" % desc.value.decode().split("|", 1)[-1].strip()
            if ps == passes[0]:
                pass
            total_fl += fl * count
            print(line, flush=True)
    print("# per-step totals (launches x count): " + " | ".join("%.2f ms  %.0f TFLOP/s" % (t, total_fl / (t * 1e-3) / 1e12 if t else 0) for t in total))


if __name__ == "__main__":
    main()
