"""fprop of the 3x3 layers with different 128-pixel box shapes (RESNET_B200_BOX=bw,bh,bn), for an ncu launch list."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from resnet_b200 import api  # noqa: E402

rng = np.random.default_rng(0)
CASES = {(56, 3, 64, 64, 1, 256): ["", "8,8,2", "16,8,1", "8,16,1", "28,4,1", "56,2,1", "14,8,1", "4,4,8", "2,2,32", "8,4,4"],
         (28, 3, 128, 128, 1, 256): ["", "4,4,8", "28,4,1", "14,8,1", "7,4,4", "14,2,4", "28,1,4", "2,2,32"],
         (14, 3, 256, 256, 1, 256): ["", "2,2,32", "14,1,9", "14,2,4", "7,2,9", "14,7,1", "1,1,128"]}
for (S, k, cin, cout, stride, N), boxes in CASES.items():
    x = rng.standard_normal((N, S, S, cin), dtype=np.float32)
    w = rng.standard_normal((cout, cin, k, k), dtype=np.float32) * 0.05
    ref = None
    for box in boxes:
        if box:
            os.environ["RESNET_B200_BOX"] = box
        else:
            os.environ.pop("RESNET_B200_BOX", None)
        y = api.conv_forward(x, w, stride, impl=0, dtype="bf16")
        ok = "" if ref is None else " same=%s" % bool(np.array_equal(y, ref))
        ref = y if ref is None else ref
        print("fprop %dx%d/%d %d->%d @%d bf16 box=%s%s" % (k, k, stride, cin, cout, S, box or "default", ok), flush=True)
