"""load_new_batch: the reference's shard format and traversal (reference: resnet.cu:1235-1325, build_training_shards.c), served by
the asynchronous prefetcher.  Two tiny shards on disk; every call must deliver exactly the bytes the reference's loader would."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_shard_traversal_and_prefetch(tmp_path):
    from resnet_b200 import api
    S, B, SHARD = 32, 4, 8
    rng = np.random.default_rng(3)
    shards = []
    for s in range(2):
        img = rng.standard_normal((SHARD, S, S, 3)).astype(np.float32)
        lab = rng.integers(0, 10, SHARD).astype(np.int32)
        img.tofile(tmp_path / ("%03d.images" % s))       # raw little-endian fp32, NHWC
        lab.tofile(tmp_path / ("%03d.labels" % s))       # raw int32
        shards.append((img, lab))
    os.environ["RESNET_B200_SHARD_DIR"] = str(tmp_path)
    try:
        t = api.Trainer(input_dim=S, n_blocks=3, reductions=[0, 1, 0], batch=B, output=10, shard_n_images=SHARD)
        bs = t.batch_struct.contents
        tr = t.t.contents
        assert bs.cur_shard_id == -1 and bs.cur_batch_in_shard == -1 and tr.cur_dump_id == -1
        expect = [(0, 0), (0, 1), (1, 0), (1, 1)]       # batches of a shard in order, then the next shard
        for i, (s, b) in enumerate(expect):
            t.load_new_batch()
            t.sync()
            img, lab = shards[s][0][b * B:(b + 1) * B], shards[s][1][b * B:(b + 1) * B]
            np.testing.assert_array_equal(api.d2h(bs.images, img.size), img.reshape(-1))
            np.testing.assert_array_equal(api.d2h(bs.correct_classes, B, np.int32), lab)
            np.testing.assert_array_equal(np.ctypeslib.as_array(bs.correct_classes_cpu, shape=(B,)), lab)
            np.testing.assert_array_equal(np.ctypeslib.as_array(bs.images_float_cpu, shape=(img.size,)), img.reshape(-1))
            assert (bs.cur_shard_id, bs.cur_batch_in_shard, tr.cur_dump_id) == (s, b + 1, i)
            if i == 1:                                    # a training step in between must not disturb the prefetched batch
                t.forward(); t.backward(); t.update()
        # epoch end: the driver resets the ids (reference: resnet.cu:3415-3416) and the traversal restarts at shard 0
        bs.cur_shard_id, bs.cur_batch_in_shard = -1, -1
        t.load_new_batch()
        t.sync()
        np.testing.assert_array_equal(api.d2h(bs.correct_classes, B, np.int32), shards[0][1][:B])
        # a missing shard is reported, not dereferenced (the reference fclose()s a NULL FILE*)
        bs.cur_shard_id, bs.cur_batch_in_shard = 1, 2
        api.L().load_new_batch(t.t, None, t.batch_struct)
        assert "cannot read batch" in api.L().resnet_b200_last_error().decode()
        api.L().resnet_b200_clear_error()
        t.close()
    finally:
        os.environ.pop("RESNET_B200_SHARD_DIR", None)
