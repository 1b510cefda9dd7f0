"""The compiled C host driver (examples/train.c) -- the reference's main() loop (reference: resnet.cu:3222-3429) written against
include/resnet.h and linked with libresnet_b200.so by plain gcc -- run as a separate process on synthetic shard files in the
reference's format: it must load batches in the reference's order, train without a recorded error, write the final dump 77777777,
and resume from it."""
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "examples", "train")


def test_c_driver_trains_on_shards_and_resumes(tmp_path):
    if not os.path.exists(EXE):
        import __graft_entry__ as g
        g.build()
    S, B, SHARD, CLASSES, BLOCKS = 32, 8, 16, 10, 3
    rng = np.random.default_rng(5)
    shard_dir, dump_root = tmp_path / "shards", tmp_path / "dumps"
    shard_dir.mkdir()
    for s in range(2):                                                  # build_training_shards.c layout: raw fp32 NHWC + int32 labels
        (rng.integers(0, 256, (SHARD, S, S, 3)).astype(np.float32) - 116.0).tofile(shard_dir / ("%03d.images" % s))
        rng.integers(0, CLASSES, SHARD).astype(np.int32).tofile(shard_dir / ("%03d.labels" % s))
    env = dict(os.environ, RESNET_B200_SHARD_DIR=str(shard_dir), RESNET_B200_DUMP_ROOT=str(dump_root), RESNET_B200_DUMP_EVERY="2")
    args = [EXE, "3", str(B), str(S), str(BLOCKS), str(CLASSES), str(SHARD)]
    r = subprocess.run(args, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    print(r.stdout[-3000:])
    assert r.returncode == 0, r.stdout[-3000:]
    assert "train ok: 3 steps" in r.stdout and "tensor cores 1" in r.stdout
    lines = [l for l in r.stdout.splitlines() if l.startswith("Epoch: 0, Batch:")]
    # traversal of the reference's loader: (shard 0, batch 0), (0, 1), (1, 0); cur_dump_id counts the calls from 0
    assert [l.split("(")[1] for l in lines] == ["shard 0, next batch 1, dump id 0)", "shard 0, next batch 2, dump id 1)", "shard 1, next batch 1, dump id 2)"]
    # the reference's checkpoint cadence (resnet.cu:2941-2944: cur_dump_id % period == 0, before the update) and the final dump
    base = dump_root / "train_c"
    assert (base / "00000000" / "model_params" / "000.buffer").exists()
    assert (base / "00000002" / "trainer_checkpoint.txt").exists()
    assert not (base / "00000001").exists()
    assert (base / "77777777" / "model_params" / "000.buffer").exists()
    # resume from the final dump: the restored cursor continues with (shard 1, batch 1)
    r2 = subprocess.run(args[:1] + ["1"] + args[2:] + ["77777777"], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    print(r2.stdout[-1500:])
    assert r2.returncode == 0 and "train ok: 1 steps" in r2.stdout, r2.stdout[-1500:]
    assert "shard 1, next batch 2" in r2.stdout
