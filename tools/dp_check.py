"""Data-parallel correctness check on real GPUs (run under torchrun, one rank per GPU; tests/test_gpu_dp.py spawns it with 2 ranks):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/dp_check.py

Checks (the reference has no data parallelism, the contract is its batch-SUM loss gradient, resnet.cu:1806-1811, and its
reverse-order update, resnet.cu:2952):
 1. dp_init makes replicas identical: every rank but 0 starts from DIFFERENT weights / Adam moments on purpose; after dp_init all
    ranks hold rank 0's bytes.
 2. the all-reduced gradient arena equals the sum of the per-shard gradients, each computed by a plain non-DP trainer on the same GPU
    (BatchNorm statistics are per shard in both computations): <= 1e-6 rel-L2, and BIT-EXACT at world 2 (a + b is commutative and the
    kernels are deterministic).
 3. after update_parameters every rank holds identical parameters, for three consecutive steps (the overlap of the bucketed
    allreduce with backward must not race with the next step's zeroing / writes).
 4. load_new_batch under DP is rank-strided: rank r receives global batches r, r + world, ... of the shard files.
"""
import ctypes as C
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from oracle import golden_cases as G          # checker-side helpers (seeded weights / batches), not the product path
    from oracle import oracle as O
    from resnet_b200 import api

    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group(backend="gloo")
    L = api.L()
    L.resnet_b200_set_device(local)
    # a well-conditioned miniature (64x64 input: no BatchNorm ever normalises over fewer than 8*8*8 values)
    cfg = dict(G.MINI4)
    cfg["batch"], cfg["input_dim"] = 8, 64
    shapes = O.param_shapes(cfg["input_dim"], cfg["n_blocks"], cfg["reductions"], output=cfg["output"])
    W = G.mini_weights(shapes)
    kw = dict(input_dim=cfg["input_dim"], n_blocks=cfg["n_blocks"], reductions=cfg["reductions"], batch=cfg["batch"], output=cfg["output"],
              lr=cfg["lr"], device=local)
    for dtype in ("tf32", "bf16"):
        t = api.Trainer(dtype=dtype, **kw)
        # rank 0 holds W; the others start from other weights and non-zero moments: dp_init must overwrite them
        t.set_params(W if rank == 0 else G.mini_weights(shapes, seed=500 + rank))
        if rank != 0:
            t.set_params([np.full_like(w, 0.25) for w in W], which=2)
            t.set_params([np.full_like(w, 0.5) for w in W], which=3)
        idbuf = (C.c_char * 128)()
        if rank == 0:
            L.resnet_b200_dp_unique_id(idbuf)
        tid = torch.tensor(list(bytes(idbuf)), dtype=torch.uint8)
        dist.broadcast(tid, src=0)
        idbuf = (C.c_char * 128)(*bytes(tid.tolist()))
        rc = L.resnet_b200_dp_init(t.t, idbuf, rank, world, 256 << 10)   # small buckets so several are in flight
        api.check()
        assert rc == 0 and L.resnet_b200_dp_world_size(t.t) == world
        for which, want in ((0, W), (2, [np.zeros_like(w) for w in W]), (3, [np.zeros_like(w) for w in W])):
            got = t.get_params(which)
            assert all(np.array_equal(g, w.reshape(-1)) for g, w in zip(got, want)), "dp_init did not broadcast tree %d" % which

        ref_t = api.Trainer(dtype=dtype, **kw)
        ref_t.set_params(W)
        worst = 0.0
        for step in range(3):
            batches = [G.mini_batch(cfg, seed=100 + 10 * step + r) for r in range(world)]
            P = t.get_params(0)
            t.set_batch(*batches[rank])
            t.forward()
            t.backward()
            got = np.concatenate(t.get_params(1))
            # reference: every shard through a plain trainer on this GPU with the same parameters, summed on the host in rank order
            ref_t.set_params(P)
            want = None
            for r in range(world):
                ref_t.set_batch(*batches[r])
                ref_t.forward()
                ref_t.backward()
                g = np.concatenate(ref_t.get_params(1))
                want = g if want is None else want + g
            err = float(np.linalg.norm(got - want) / np.linalg.norm(want))
            worst = max(worst, err)
            assert err <= 1e-6, (dtype, step, err)
            if world == 2:
                assert np.array_equal(got, want), (dtype, step, "all-reduced gradient is not bit-identical to a + b")
            t.update()
            p = np.concatenate(t.get_params(0))
            ps = [torch.zeros(p.size, dtype=torch.float32) for _ in range(world)]
            dist.all_gather(ps, torch.from_numpy(p))
            assert all(torch.equal(ps[0], q) for q in ps), (dtype, step, "parameters differ across ranks")
        print("rank %d %s: allreduced gradient == sum of shard gradients (worst rel-L2 %.1e%s), parameters identical across ranks for 3 steps"
              % (rank, dtype, worst, ", bit-exact" if world == 2 else ""), flush=True)
        t.close()
        ref_t.close()

    # ---- rank-strided load_new_batch: one shard directory, every rank walks the same global sequence at its own offset
    S, B, SHARD, NSH = 32, 4, 8, 3
    shard_dir = os.path.join(tempfile.gettempdir(), "resnet_b200_dp_shards")
    rng = np.random.default_rng(3)
    shards = []
    for s in range(NSH):
        img = rng.standard_normal((SHARD, S, S, 3)).astype(np.float32)
        lab = rng.integers(0, 10, SHARD).astype(np.int32)
        shards.append((img, lab))
    if rank == 0:
        os.makedirs(shard_dir, exist_ok=True)
        for s, (img, lab) in enumerate(shards):
            img.tofile(os.path.join(shard_dir, "%03d.images" % s))
            lab.tofile(os.path.join(shard_dir, "%03d.labels" % s))
    dist.barrier()
    os.environ["RESNET_B200_SHARD_DIR"] = shard_dir
    t = api.Trainer(input_dim=S, n_blocks=3, reductions=[0, 1, 0], batch=B, output=10, shard_n_images=SHARD, device=local)
    idbuf = (C.c_char * 128)()
    if rank == 0:
        L.resnet_b200_dp_unique_id(idbuf)
    tid = torch.tensor(list(bytes(idbuf)), dtype=torch.uint8)
    dist.broadcast(tid, src=0)
    idbuf = (C.c_char * 128)(*bytes(tid.tolist()))
    assert L.resnet_b200_dp_init(t.t, idbuf, rank, world, 0) == 0
    seq = [(s, b) for s in range(NSH) for b in range(SHARD // B)]          # the single-GPU traversal
    bs = t.batch_struct.contents
    for k in range(len(seq) // world):
        s, b = seq[k * world + rank]
        t.load_new_batch()
        t.sync()
        np.testing.assert_array_equal(api.d2h(bs.correct_classes, B, np.int32), shards[s][1][b * B:(b + 1) * B])
        np.testing.assert_array_equal(api.d2h(bs.images, B * S * S * 3), shards[s][0][b * B:(b + 1) * B].reshape(-1))
    t.close()
    dist.barrier()
    if rank == 0:
        print("dp_check ok (world %d): broadcast at init, gradient equality, identical parameters, rank-strided loader" % world, flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
