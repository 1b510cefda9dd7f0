"""Where does a conv tile's time go?  fprop of three layer shapes with (a) everything, (b) no MMAs issued (TMA feed + epilogue),
(c) no MMAs and no epilogue work (TMA feed only), and with fewer pipeline stages -- for an ncu launch list."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from resnet_b200 import api  # noqa: E402

rng = np.random.default_rng(0)
for (S, k, cin, cout, stride, N) in [(56, 3, 64, 64, 1, 256), (28, 3, 128, 128, 1, 256), (28, 3, 512, 1024, 2, 256)]:
    x = rng.standard_normal((N, S, S, cin), dtype=np.float32)
    w = rng.standard_normal((cout, cin, k, k), dtype=np.float32) * 0.05
    for skip, stages in [("0", ""), ("1", ""), ("3", ""), ("0", "4"), ("3", "4"), ("0", "2"), ("3", "2")]:
        os.environ["RESNET_B200_DEBUG_SKIP"] = skip
        if stages:
            os.environ["RESNET_B200_STAGES"] = stages
        else:
            os.environ.pop("RESNET_B200_STAGES", None)
        api.conv_forward(x, w, stride, impl=0, dtype="bf16")
        print("fprop %dx%d/%d %d->%d @%d bf16 skip=%s stages=%s" % (k, k, stride, cin, cout, S, skip, stages or "max"), flush=True)
