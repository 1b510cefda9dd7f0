"""Two ResNet-50 training steps at batch 256 (or --batch) through the C API and nothing else: the smallest program that shows
every kernel of the hot path, for `ncu` captures (profiles/)."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from resnet_b200 import synth as O  # noqa: E402  (synthetic batch generator)
from resnet_b200 import api  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--steps", type=int, default=2)
ap.add_argument("--dtype", default="tf32", choices=["tf32", "bf16"])
ap.add_argument("--blocks", type=int, default=16, choices=[16, 50], help="16 = ResNet-50, 50 = ResNet-152")
a = ap.parse_args()
red = [1 if i in ((3, 7, 13) if a.blocks == 16 else (3, 11, 47)) else 0 for i in range(a.blocks)]
t = api.Trainer(input_dim=224, n_blocks=a.blocks, reductions=red, batch=a.batch, output=1000, lr=1e-4, seed=1234, device=0, dtype=a.dtype)
img, lab = O.synthetic_batch(a.batch, 224, seed=1234)
for _ in range(a.steps):
    t.set_batch(img, lab)
    t.forward()
    t.backward()
    t.update()
t.sync()
print("one_step ok, launches", api.L().resnet_b200_launch_count())
