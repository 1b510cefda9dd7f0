"""A handful of full-size single operators through the C API, for ONE `ncu --set full --import-source on` capture of the kernels
that dominate the step (profiles/): the best and the worst convolution shapes in bf16 and TF32, their wgrad, and the BatchNorm
kernels on a tensor larger than L2.  Prints one line per operator so the capture's launch order can be read back."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from resnet_b200 import api  # noqa: E402

N = int(os.environ.get("PROBE_BATCH", "256"))
rng = np.random.default_rng(0)
CONVS = [(28, 3, 512, 1024, 2), (56, 3, 64, 64, 1), (56, 1, 64, 256, 1)]
for dtype in ("bf16", "f32"):
    for (S, k, cin, cout, stride) in CONVS:
        x = rng.standard_normal((N, S, S, cin), dtype=np.float32)
        w = rng.standard_normal((cout, cin, k, k), dtype=np.float32) * 0.05
        dy = rng.standard_normal((N, S // stride, S // stride, cout), dtype=np.float32)
        os.environ["RESNET_B200_OP_STATS"] = "1"
        y = api.conv_forward(x, w, stride, impl=0, dtype=dtype)
        os.environ.pop("RESNET_B200_OP_STATS")
        din, dw = api.conv_backward(x, w, dy, stride, impl=0, dtype=dtype)
        print("conv %dx%d/%d %d->%d @%d N=%d %s: fprop(+stats), dgrad, wgrad, wgrad_reduce  |y|=%.3f |dw|=%.3f" %
              (k, k, stride, cin, cout, S, N, dtype, float(np.abs(y).mean()), float(np.abs(dw).mean())), flush=True)
    x = rng.standard_normal((N, 28, 28, 512), dtype=np.float32)
    g = np.ones(512, np.float32)
    b = np.zeros(512, np.float32)
    dy = rng.standard_normal(x.shape, dtype=np.float32)
    mu, var, y = api.batchnorm_forward(x, g, b, 1e-7, True, dtype=dtype)
    dg, db, dx = api.batchnorm_backward(x, g, 1e-7, mu, var, y, dy, True, dtype=dtype)
    print("batchnorm rows=%d C=512 %s: bn_reduce<fwd>, bn_finalize, bn_apply, bn_reduce<bwd>, bn_bwd_finalize, bn_bwd_dx  |dx|=%.3f" %
          (N * 28 * 28, dtype, float(np.abs(dx).mean())), flush=True)
