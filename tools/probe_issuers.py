"""One, two or four MMA-issuing threads per CTA (RESNET_B200_ISSUERS) on full-size layers: fprop, dgrad + wgrad of each shape, one
igemm launch per printed line (fprop: 1 line; backward: dgrad then wgrad), for an ncu launch list (-k regex:igemm)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from resnet_b200 import api  # noqa: E402

# (the first run of this tool at batch 256 with six shapes x two dtypes did not finish in 5 GPU-minutes: the host side of api.conv_* -- numpy
# conversions and PCIe copies of 100-400 MB tensors per call -- dominates; batch 128 and four shapes keep every CTA at >= 2 tiles)
SHAPES = [(28, 3, 512, 1024, 2), (14, 3, 256, 256, 1), (28, 3, 128, 128, 1), (56, 3, 64, 64, 1)]
N = 128
rng = np.random.default_rng(0)
dtypes = sys.argv[1:] or ["bf16"]
for (S, k, cin, cout, stride) in SHAPES:
    x = rng.standard_normal((N, S, S, cin), dtype=np.float32)
    w = rng.standard_normal((cout, cin, k, k), dtype=np.float32) * 0.05
    dy = rng.standard_normal((N, S // stride, S // stride, cout), dtype=np.float32)
    for dtype in dtypes:
        variants = [("1", "0"), ("2", "0"), ("4", "0")]
        if k == 3 and stride == 1 and cout <= 128:
            variants += [("2", "1"), ("4", "1")]
        for iss, halo in variants:
            os.environ["RESNET_B200_ISSUERS"] = iss
            os.environ["RESNET_B200_HALO"] = halo
            y = api.conv_forward(x, w, stride, impl=0, dtype=dtype)
            tag = "%dx%d/%d %d->%d @%d %s issuers=%s halo=%s" % (k, k, stride, cin, cout, S, dtype, iss, halo)
            print("fprop " + tag + " mean|y|=%.4f" % float(np.abs(y).mean()), flush=True)
            din, dw = api.conv_backward(x, w, dy, stride, impl=0, dtype=dtype)
            print("dgrad " + tag + " mean|dx|=%.4f" % float(np.abs(din).mean()), flush=True)
            print("wgrad " + tag + " mean|dw|=%.4f" % float(np.abs(dw).mean()), flush=True)
