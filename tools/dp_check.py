"""Data-parallel correctness check on real GPUs (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/dp_check.py

Every rank runs forward + backward on ITS shard with the NCCL bucketed allreduce enabled, then rank r also runs every
shard through a second, non-DP trainer and sums the gradients itself.  The all-reduced gradient arena must equal that
sum (the loss gradient is a batch SUM, reference: resnet.cu:1806-1811, so allreduce-SUM == large-batch gradient; BatchNorm
statistics are per shard in both computations).  After update_parameters all ranks must hold identical parameters.
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from oracle import golden_cases as G
    from oracle import oracle as O
    from resnet_b200 import api

    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group(backend="gloo")
    L = api.L()
    L.resnet_b200_set_device(local)
    # a well-conditioned miniature (64x64 input: no BatchNorm ever normalises over fewer than 8*8*8 values); tiny 1x1-spatial nets
    # amplify the last-bit nondeterminism of the BatchNorm block sums by 1e3 and cannot resolve a 1e-3 bar
    cfg = dict(G.MINI4)
    cfg["batch"], cfg["input_dim"] = 8, 64
    shapes = O.param_shapes(cfg["input_dim"], cfg["n_blocks"], cfg["reductions"], output=cfg["output"])
    W = G.mini_weights(shapes)
    kw = dict(input_dim=cfg["input_dim"], n_blocks=cfg["n_blocks"], reductions=cfg["reductions"], batch=cfg["batch"], output=cfg["output"],
              lr=cfg["lr"], device=local)
    t = api.Trainer(**kw)
    t.set_params(W)
    idbuf = (C.c_char * 128)()
    if rank == 0:
        L.resnet_b200_dp_unique_id(idbuf)
    tid = torch.tensor(list(bytes(idbuf)), dtype=torch.uint8)
    dist.broadcast(tid, src=0)
    idbuf = (C.c_char * 128)(*bytes(tid.tolist()))
    rc = L.resnet_b200_dp_init(t.t, idbuf, rank, world, 4 << 20)   # small buckets so several are in flight
    api.check()
    assert rc == 0 and L.resnet_b200_dp_world_size(t.t) == world
    batches = [G.mini_batch(cfg, seed=100 + r) for r in range(world)]
    t.set_batch(*batches[rank])
    t.forward()
    t.backward()
    got = np.concatenate(t.get_params(1))
    # reference: every shard through a plain trainer on this GPU, summed on the host
    ref_t = api.Trainer(**kw)
    ref_t.set_params(W)
    want = None
    for r in range(world):
        ref_t.set_batch(*batches[r])
        ref_t.forward()
        ref_t.backward()
        g = np.concatenate(ref_t.get_params(1))
        want = g if want is None else want + g
        ref_t.update()             # zeroes the gradients (and moves weights: reset them)
        ref_t.set_params(W)
        ref_t.set_params([np.zeros_like(w) for w in W], which=2)
        ref_t.set_params([np.zeros_like(w) for w in W], which=3)
    err = float(np.linalg.norm(got - want) / np.linalg.norm(want))
    t.update()
    p = np.concatenate(t.get_params(0))
    ps = [torch.zeros(p.size, dtype=torch.float32) for _ in range(world)]
    dist.all_gather(ps, torch.from_numpy(p))
    same = all(torch.equal(ps[0], q) for q in ps)
    print("rank %d: allreduced-gradient rel-L2 vs summed shards %.3e ; parameters identical across ranks: %s" % (rank, err, same), flush=True)
    assert err < 2e-3, err
    assert same
    dist.barrier()
    if rank == 0:
        print("dp_check ok (world %d)" % world, flush=True)


if __name__ == "__main__":
    main()
