"""world_size-2 gloo tests (CPU) of the host-side multi-GPU plumbing in bench.py: rendezvous on 127.0.0.1, broadcast of the
128-byte NCCL id from rank 0, barrier, max / sum over ranks of the per-rank timings, per-rank synthetic shards."""
import os
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    os.environ.update({"RANK": str(rank), "LOCAL_RANK": str(rank), "WORLD_SIZE": str(world), "MASTER_ADDR": "127.0.0.1",
                       "MASTER_PORT": str(port)})
    sys.path.insert(0, ROOT)
    import bench
    from oracle import oracle as O
    r, lr, w = bench.dist_env()
    assert (r, lr, w) == (rank, rank, world)
    dist = bench.init_dist(world)
    payload = bytes(range(128)) if rank == 0 else bytes(128)
    got = bench.broadcast_bytes(dist, payload, 128)
    bench.barrier(dist)
    mx = bench.reduce_max(dist, 10.0 + rank)
    sm = bench.reduce_sum(dist, 1.0 + rank)
    img, lab = O.synthetic_batch(2, 8, seed=1234 + 1000 * rank)       # the per-rank shard rule bench.py uses
    q.put((rank, got == bytes(range(128)), mx, sm, float(img.sum()), lab.tolist()))
    bench.barrier(dist)
    dist.destroy_process_group()


def test_gloo_world2_plumbing():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29400 + os.getpid() % 500
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, *_ in res)                 # every rank holds rank 0's id bytes
    assert all(mx == 11.0 and sm == 3.0 for _, _, mx, sm, *_ in res)   # max / sum over ranks
    assert res[0][4] != res[1][4]                       # ranks train on different shards


def test_single_process_helpers_are_identity():
    sys.path.insert(0, ROOT)
    import bench
    assert bench.reduce_max(None, 3.5) == 3.5 and bench.reduce_sum(None, 2.0) == 2.0
    assert bench.broadcast_bytes(None, b"abc", 3) == b"abc"
    p = bench.peaks()
    assert p["hbm_gbs"] > 1000 and p["bf16_tflops_sustained"] > 100
    assert abs(bench.FLOP_PER_IMAGE_STEP * 256 - 10.007e12) < 1e10   # 10.01 TFLOP per batch-256 step
