"""Runs single convolutions of full ResNet-50 layer shapes through the C API under different epilogue settings, for an
`ncu --metrics gpu__time_duration.sum` launch list: one igemm launch per printed line, in order.
    python tools/conv_probe.py > probe.log ; ncu ... python tools/conv_probe.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from resnet_b200 import api  # noqa: E402

SHAPES = [(56, 1, 64, 256, 1, 256), (28, 1, 128, 512, 1, 256), (56, 1, 256, 64, 1, 256), (56, 3, 64, 64, 1, 256), (14, 3, 256, 256, 1, 256)]
VARIANTS = [dict(), dict(RESNET_B200_OP_STATS="1"), dict(RESNET_B200_OP_STATS="1", RESNET_B200_EPI_GROUPS="1"),
            dict(RESNET_B200_EPI_GROUPS="1"), dict(RESNET_B200_EPI_GROUPS="1", RESNET_B200_NSTAGING="4")]
rng = np.random.default_rng(0)
for (S, k, cin, cout, stride, N) in SHAPES:
    x = rng.standard_normal((N, S, S, cin), dtype=np.float32)
    w = (rng.standard_normal((cout, cin, k, k), dtype=np.float32) * 0.05)
    for dtype in ("bf16", "f32"):
        for v in VARIANTS:
            for kk in ("RESNET_B200_OP_STATS", "RESNET_B200_EPI_GROUPS", "RESNET_B200_NSTAGING"):
                os.environ.pop(kk, None)
            os.environ.update(v)
            y = api.conv_forward(x, w, stride, impl=0, dtype=dtype)
            print("fprop %dx%d/%d %d->%d @%d N=%d %s %s  mean|y|=%.4f" % (k, k, stride, cin, cout, S, N, dtype, v, float(np.abs(y).mean())), flush=True)
