import os, sys
import numpy as np
sys.path.insert(0, "/root/repo" if os.path.isdir("/root/repo/resnet_b200") else os.getcwd())
from resnet_b200 import api
from oracle import oracle as O
rng = np.random.default_rng(0)
SH = [(56, 3, 64, 64, 1, 256), (56, 1, 64, 256, 1, 256), (56, 1, 256, 128, 1, 256), (224, 7, 3, 64, 2, 256)]
for (S, k, cin, cout, stride, N) in SH:
    if cin == 3:
        x, _ = O.synthetic_batch(N, S, seed=1)
    else:
        x = rng.standard_normal((N, S, S, cin), dtype=np.float32)
    w = (rng.standard_normal((cout, cin, k, k), dtype=np.float32) * 0.05)
    for dtype in ("bf16", "f32"):
        for res in ("1", "0"):
            os.environ["RESNET_B200_RESIDENT_B"] = res
            y = api.conv_forward(x, w, stride, impl=0, dtype=dtype)
            print("fprop %dx%d/%d %d->%d @%d %s residentB=%s |y|=%.4f" % (k, k, stride, cin, cout, S, dtype, res, float(np.abs(y).mean())), flush=True)
