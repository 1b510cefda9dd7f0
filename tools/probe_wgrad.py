"""wgrad of the small-channel layers (several taps per tile) with the taps issued as one wide MMA vs one MMA per tap, for an ncu
launch list (igemm_mnmajor launches only; one per printed line)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
from resnet_b200 import api  # noqa: E402

rng = np.random.default_rng(0)
for (S, k, cin, cout, stride, N) in [(56, 3, 64, 64, 1, 256), (28, 3, 128, 128, 1, 256), (224, 7, 3, 64, 2, 256)]:
    x = O.synthetic_batch(N, S, seed=1)[0] if cin == 3 else rng.standard_normal((N, S, S, cin), dtype=np.float32)
    w = rng.standard_normal((cout, cin, k, k), dtype=np.float32) * 0.05
    dy = rng.standard_normal((N, S // stride, S // stride, cout), dtype=np.float32)
    for dtype in ("bf16", "f32"):
        ref = None
        for merge in ("1", "0"):
            os.environ["RESNET_B200_WGRAD_MERGE"] = merge
            _, dw = api.conv_backward(x, w, dy, stride, want_din=False, impl=0, dtype=dtype)
            same = "" if ref is None else " identical=%s" % bool((dw == ref).all())
            ref = dw
            print("wgrad %dx%d/%d %d->%d @%d %s merge=%s |dw|=%.4f%s" % (k, k, stride, cin, cout, S, dtype, merge, float(np.abs(dw).mean()), same), flush=True)
