"""Synthetic input generator shared by bench.py, the tests and the oracle (pure numpy, no device code).

SURVEY.md 8(d): images are i.i.d. uniform integers 0..255 minus the per-channel means the reference's shard builder subtracts
(reference: build_training_shards.c:120-134), fp32 NHWC; labels uniform in [0, n_classes).  Lives in the package -- not under
oracle/ -- so that the benchmark's product arm never imports the checker."""
import numpy as np

CHANNEL_MEANS = (103.94, 116.78, 123.68)


def synthetic_batch(batch, input_dim, seed=1234, n_classes=1000):
    rng = np.random.default_rng(seed)
    img = rng.integers(0, 256, size=(batch, input_dim, input_dim, 3)).astype(np.float32)
    img -= np.array(CHANNEL_MEANS, np.float32)
    labels = np.random.default_rng(seed + 3087).integers(0, n_classes, size=batch).astype(np.int32)
    return np.ascontiguousarray(img), labels
