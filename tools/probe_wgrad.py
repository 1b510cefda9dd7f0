"""wgrad A/B under ncu (igemm_mnmajor launches only; one per printed line): PROBE=merge -- the taps of a tile issued as one wide MMA
vs one MMA per tap (small-channel layers); PROBE=mpair -- two 128-row co tiles per work item sharing the X tile vs one."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
from resnet_b200 import api  # noqa: E402

which = os.environ.get("PROBE", "merge")
var = {"merge": "RESNET_B200_WGRAD_MERGE", "mpair": "RESNET_B200_WGRAD_MPAIR"}[which]
shapes = {"merge": [(56, 3, 64, 64, 1, 256), (28, 3, 128, 128, 1, 256), (224, 7, 3, 64, 2, 256)],
          "mpair": [(56, 3, 256, 512, 2, 256), (28, 3, 512, 1024, 2, 128), (14, 3, 256, 256, 1, 256), (14, 1, 1024, 256, 1, 256), (7, 1, 512, 2048, 1, 256)]}[which]
rng = np.random.default_rng(0)
for (S, k, cin, cout, stride, N) in shapes:
    x = O.synthetic_batch(N, S, seed=1)[0] if cin == 3 else rng.standard_normal((N, S, S, cin), dtype=np.float32)
    w = rng.standard_normal((cout, cin, k, k), dtype=np.float32) * 0.05
    dy = rng.standard_normal((N, S // stride, S // stride, cout), dtype=np.float32)
    for dtype in ("bf16", "f32"):
        ref = None
        for on in ("1", "0"):
            os.environ[var] = on
            _, dw = api.conv_backward(x, w, dy, stride, want_din=False, impl=0, dtype=dtype)
            same = "" if ref is None else " max|diff|/max=%.1e" % float(np.abs(dw - ref).max() / np.abs(ref).max())
            ref = dw
            print("wgrad %dx%d/%d %d->%d @%d N=%d %s %s=%s%s" % (k, k, stride, cin, cout, S, N, dtype, which, on, same), flush=True)
