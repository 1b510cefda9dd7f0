"""Bring-up probe for the tcgen05 conv kernels: compares the tensor-core path with the fp32 SIMT path on device for
each geometry, one subprocess per case (a hung kernel must not take the whole GPU session down).
usage: python tools/tc_probe.py [out.txt]"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = [
    # S, k, cin, cout, stride, N
    (8, 1, 64, 64, 1, 2), (8, 3, 64, 64, 1, 2), (8, 3, 128, 128, 2, 2), (8, 3, 256, 512, 2, 2), (14, 3, 256, 256, 1, 4),
    (7, 3, 512, 512, 1, 2), (7, 1, 512, 2048, 1, 4), (56, 3, 64, 64, 1, 2), (56, 1, 256, 128, 1, 2), (28, 3, 128, 128, 1, 3),
    (14, 3, 512, 512, 2, 2), (56, 3, 256, 512, 2, 8), (14, 3, 1024, 2048, 2, 8),
]
# wgrad descriptor candidates "lbo,sbo,layout_type,tma_swizzle" tried only if the built-in default is wrong
WGRAD_VARIANTS = ["4096,512,1,4", "512,4096,1,4", "4096,1024,1,4", "4096,1024,2,3", "1024,4096,2,3"]

CHILD = r"""
import sys, numpy as np
sys.path.insert(0, %r)
from resnet_b200 import api
S,k,cin,cout,stride,N = %r
wgrad_only = %r
rng = np.random.default_rng(0)
x = rng.standard_normal((N,S,S,cin)).astype(np.float32)
w = (rng.standard_normal((cout,cin,k,k))*0.1).astype(np.float32)
dy = rng.standard_normal((N,S//stride,S//stride,cout)).astype(np.float32)
base = rng.standard_normal(x.shape).astype(np.float32)
def rel(a,b):
    return float(np.abs(a-b).max()/max(1e-9,np.abs(b).max()))
out=[]
din1,dw1 = api.conv_backward(x,w,dy,stride,impl=1)
if not wgrad_only:
    y1 = api.conv_forward(x,w,stride,impl=1)
    try:
        y0 = api.conv_forward(x,w,stride,impl=0); out.append('fprop %%.2e'%%rel(y0,y1))
    except Exception as e: out.append('fprop ERR %%s'%%e)
try:
    din0,dw0 = api.conv_backward(x,w,dy,stride,impl=0); out.append('dgrad %%.2e wgrad %%.2e (|dw0| %%.3g |dw1| %%.3g)'%%(rel(din0,din1),rel(dw0,dw1),np.abs(dw0).max(),np.abs(dw1).max()))
    if not wgrad_only:
        dina,_ = api.conv_backward(x,w,dy,stride,din_base=base,impl=0); out.append('dgrad_add %%.2e'%%rel(dina,base+din1))
except Exception as e: out.append('bwd ERR %%s'%%e)
print(' | '.join(out))
"""


def run(case, wgrad_only=False, desc=None):
    env = dict(os.environ)
    if desc:
        env["RESNET_B200_WGRAD_DESC"] = desc
    try:
        r = subprocess.run([sys.executable, "-c", CHILD % (ROOT, case, wgrad_only)], capture_output=True, text=True, timeout=60, env=env)
        msg = (r.stdout.strip().splitlines() or ["<no output>"])[-1]
        if r.returncode != 0:
            msg += " | rc=%d %s" % (r.returncode, r.stderr.strip()[-300:])
    except subprocess.TimeoutExpired:
        msg = "TIMEOUT (hang)"
    return msg


def main():
    out_path = sys.argv[1] if len(sys.argv) > 1 else None
    lines = []

    def emit(line):
        print(line, flush=True)
        lines.append(line)
        if out_path:
            open(out_path, "w").write("\n".join(lines) + "\n")

    first = run(CASES[0])
    emit("%s: %s" % (CASES[0], first))
    wgrad_bad = "wgrad" not in first or float(first.split("wgrad ")[1].split()[0]) > 1e-2
    if wgrad_bad:
        for v in WGRAD_VARIANTS:
            for case in (CASES[0], CASES[1]):
                emit("variant %s %s: %s" % (v, case, run(case, wgrad_only=True, desc=v)))
    for case in CASES[1:]:
        emit("%s: %s" % (case, run(case)))


if __name__ == "__main__":
    main()
