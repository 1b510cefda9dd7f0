"""CPU tests of the host-side helpers around the C ABI (no GPU, no compute calls)."""
import numpy as np


def test_bf16_helpers_match_torch_rounding():
    """api.to_bf16 / from_bf16 are what the bf16 tests feed the oracle with: round-to-nearest-even, as cvt.rn.bf16.f32 and
    torch.bfloat16 do, including ties, subnormal neighbours, large values and signed zeros."""
    import torch
    from resnet_b200 import api
    rng = np.random.default_rng(0)
    a = np.concatenate([rng.standard_normal(200000).astype(np.float32) * 100, rng.standard_normal(1000).astype(np.float32) * 1e-30,
                        np.array([0.0, -0.0, 1.0, 1.00390625, 1.01171875, 3.3895314e38, -65504.0, 1e-38], np.float32)])
    ours = api.from_bf16(api.to_bf16(a))
    ref = torch.from_numpy(a).to(torch.bfloat16).float().numpy()
    np.testing.assert_array_equal(ours.view(np.uint32), ref.view(np.uint32))
    np.testing.assert_array_equal(api.bf16_round(a.reshape(-1, 8)).reshape(-1), ours)
    assert api.to_bf16(a).dtype == np.uint16


def test_bench_configs_match_baseline_json():
    """bench.py --config names the BASELINE.json configurations: batch sizes, dtypes and the FLOP model of BASELINE.md section 3."""
    import json
    import os
    import bench
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cfgs = json.load(open(os.path.join(root, "BASELINE.json")))["configs"]
    assert "batch 256" in cfgs[1] and "TF32" in cfgs[1] and bench.CONFIGS["c2"]["batch"] == 256 and bench.CONFIGS["c2"]["dtype"] == "tf32"
    assert "batch 1024" in cfgs[2] and "bf16" in cfgs[2] and bench.CONFIGS["c3"]["batch"] == 1024 and bench.CONFIGS["c3"]["fwd_only"]
    assert "batch 256/GPU" in cfgs[3] and bench.CONFIGS["c4"]["dtype"] == "bf16" and not bench.CONFIGS["c4"]["fwd_only"]
    assert "ResNet-152" in cfgs[4] and "batch 128/GPU" in cfgs[4] and bench.CONFIGS["c5"]["batch"] == 128 and len(bench.CONFIGS["c5"]["red"]) == 50
    # FLOP model: 2 N Ho Wo Cout Cin k^2 per pass; training = 3 x forward - stem dgrad
    from oracle import oracle as O
    def conv_gflop(n_blocks, red):
        f = 2 * 112 * 112 * 64 * 3 * 49
        stem = f
        for b in O.block_plan(224, n_blocks, red):
            S, So = b["spatial"], b["spatial"] // b["stride"]
            f += 2 * S * S * b["reduced"] * b["incoming"] + 2 * So * So * b["reduced"] * b["reduced"] * 9 + 2 * So * So * b["expanded"] * b["reduced"]
            if b["proj"]:
                f += 2 * So * So * b["expanded"] * b["incoming"] * b["proj_k"] ** 2
        return f / 1e9, stem / 1e9
    f50, stem = conv_gflop(16, bench.R50_REDUCTIONS)
    assert abs(f50 + 0.0041 - bench.CONFIGS["c3"]["gflop"]) < 0.02
    assert abs(3 * (f50 + 0.0041) - stem - bench.CONFIGS["c2"]["gflop"]) < 0.05
    f152, _ = conv_gflop(50, bench.R152_REDUCTIONS)
    assert abs(3 * (f152 + 0.0041) - stem - bench.CONFIGS["c5"]["gflop"]) < 0.1


def test_loader_traversal_single_and_rank_strided():
    """load_new_batch's cursor logic without a GPU (resnet_b200_loader_plan): one rank walks the reference's sequence (batches of a
    shard in order, then the next shard; reference: resnet.cu:1260-1295); under data parallelism rank r of `world` takes positions
    r, r + world, ... of that same sequence, so the ranks' deliveries are disjoint and their union, in order, is the single-GPU run."""
    import ctypes as C
    from resnet_b200 import lib
    L = lib.load()

    def plan(rank, world, batch, shard_n, n):
        s, b = (C.c_int * n)(), (C.c_int * n)()
        assert L.resnet_b200_loader_plan(rank, world, batch, shard_n, n, s, b) == 0
        return list(zip(list(s), list(b)))
    single = plan(0, 1, 4, 12, 9)
    assert single == [(0, 0), (0, 1), (0, 2), (1, 0), (1, 1), (1, 2), (2, 0), (2, 1), (2, 2)]
    for world in (2, 3, 4):
        per_rank = [plan(r, world, 4, 12, 6) for r in range(world)]
        merged = [per_rank[i % world][i // world] for i in range(6 * world)]
        assert merged == plan(0, 1, 4, 12, 6 * world), (world, merged)
    # a shard size that is not a multiple of the batch: the reference moves on when batch * batch_size >= shard_n_images
    assert plan(0, 1, 5, 12, 5) == [(0, 0), (0, 1), (0, 2), (1, 0), (1, 1)]


def test_bench_arms_share_one_config_dict():
    """the product arm and the reference arms describe the workload with the same `config` dictionary (the driver compares them)"""
    import bench
    for name, cfg in bench.CONFIGS.items():
        a, b = bench.config_dict(name, cfg["batch"], 1), bench.config_dict(name, cfg["batch"], 1)
        assert a == b and a["baseline_config"] == name and a["global_batch"] == cfg["batch"]
        assert bench.config_dict(name, cfg["batch"], 8)["global_batch"] == 8 * cfg["batch"]
    assert bench.metric_of("c2") == bench.metric_of("c4") == bench.METRIC and "forward" in bench.metric_of("c3")


def test_bench_headline_roofline_family():
    """bench.py pick_dominant: the largest family wins, the sub-family lines of the tcgen05 kernel do not compete, and a near tie
    (within 5 %) with the tcgen05 fprop + dgrad family is resolved in its favour so that the headline does not flip from run to run."""
    import bench
    km = {"kernel": "igemm_kmajor_kernel (tcgen05 fprop+dgrad)", "ms_per_step": 16.5}
    sub = {"kernel": "igemm_kmajor_kernel, 3x3 + stem fprop/dgrad", "ms_per_step": 30.0}
    wg = {"kernel": "igemm_mnmajor_kernel (tcgen05 wgrad + split-K reduce)", "ms_per_step": 8.4}
    bn = {"kernel": "BatchNorm/elementwise", "ms_per_step": 16.7}
    assert bench.pick_dominant([km, sub, wg, bn]) is km          # 1 % behind: tie, stays on the tensor-core family; `sub` never competes
    assert bench.pick_dominant([km, wg, dict(bn, ms_per_step=18.0)])["kernel"] == "BatchNorm/elementwise"   # a clear lead wins
    assert bench.pick_dominant([wg, bn]) is bn                   # forward-only style list without the family
    assert bench.pick_dominant([dict(km, ms_per_step=20.0), bn])["kernel"].startswith("igemm_kmajor_kernel")
