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


def test_multi_issuer_barrier_protocol_model():
    """The mbarrier protocol of the (experimental) multi-issuer convolution kernels, model-checked under random schedules with TMA loads
    completing out of order (tools/issuer_protocol_sim.py).  Ownership by ring slot -- what igemm.cu does -- never deadlocks, never lets
    an issuer onto a stale slot, never touches an accumulator another tile owns.  Ownership by stage index -- the first version -- fails
    on 3-stage rings and on work items with an odd stage count, the two places where the batch-256 step hung / faulted on the GPU
    (profiles/r01_issuers_status.txt)."""
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location("issuer_protocol_sim", os.path.join(os.path.dirname(__file__), "..", "tools", "issuer_protocol_sim.py"))
    sim = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(sim)
    for cfg, failing in sim.sweep(seeds=10, tiles=6).items():
        assert failing == 0, ("by slot", cfg, failing)
    old = sim.sweep(seeds=20, tiles=6, by_slot=False, configs=[(3, 8, 2), (4, 339, 2), (4, 8, 2)])
    assert old[(3, 8, 2)] > 0 and old[(4, 339, 2)] > 0, old      # the hardware failures, reproduced
    assert old[(4, 8, 2)] == 0, old                               # ... and the configuration that ran
    for cfg, failing in sim.sweep_halo(seeds=8, tiles=4).items():   # the haloed-patch kernel's two rings
        assert failing == 0, ("halo", cfg, failing)


def _halo_emulation(x, wk, tdx, tdy, bw, bh):
    """numpy restatement of igemm_halo_kernel's addressing: x [S][S][K]; wk [ncol][tap][K]; one (bh + 2) x (bw + 2) patch per tile laid out
    as rows of pitch PW; accumulator row r reads patch row r + (dy + 1) * PW + (dx + 1) for tap (dx, dy); rows past the patch are NaN."""
    import numpy as np
    S, _, K = x.shape
    ncol = wk.shape[0]
    PW, PH = bw + 2, bh + 2
    out = np.zeros((S, S, ncol))
    for oh0 in range(0, S, bh):
        for ow0 in range(0, S, bw):
            patch = np.full((2 * PW + 2 + 128, K), np.nan)
            for h in range(PH):
                for w in range(PW):
                    gh, gw = oh0 - 1 + h, ow0 - 1 + w
                    patch[h * PW + w] = x[gh, gw] if (0 <= gh < S and 0 <= gw < S) else 0.0   # TMA zero fill = the padding
            acc = np.zeros((128, ncol))
            for t in range(9):
                off = (tdy[t] + 1) * PW + (tdx[t] + 1)
                acc += np.nan_to_num(patch[off:off + 128], nan=1e30) @ wk[:, t, :].T      # garbage rows must only reach dropped outputs
            for pos in range(128):
                h, w = divmod(pos, PW)
                if w < bw and h < bh and oh0 + h < S and ow0 + w < S:
                    out[oh0 + h, ow0 + w] = acc[pos]
    return out


def test_halo_patch_addressing_emulation():
    """The haloed-patch addressing of igemm_halo_kernel (one patch per tile, taps = start offsets of whole rows, useful rows
    pos = h * PW + w with w < bw, h < bh) restated in numpy against a direct 3x3 convolution and its transpose, for the two tile
    shapes ResNet-50 uses (14 x 8 at 56 x 56, 28 x 4 at 28 x 28 -- here on smaller maps) and ragged maps; fprop taps read
    (h + kh - 1, w + kw - 1), dgrad taps (h + 1 - kh, w + 1 - kw) (igemm.cu make_halo_plan callers)."""
    import numpy as np
    rng = np.random.default_rng(0)

    def conv_ref(x, w):
        S, co = x.shape[0], w.shape[0]
        xp = np.pad(x, ((1, 1), (1, 1), (0, 0)))
        y = np.zeros((S, S, co))
        for kh in range(3):
            for kw in range(3):
                y += xp[kh:kh + S, kw:kw + S, :] @ w[:, :, kh, kw].T
        return y
    for S, bw, bh in [(14, 14, 7), (28, 28, 4), (20, 14, 8), (9, 9, 9)]:
        assert bh * (bw + 2) <= 128
        ci, co = 5, 6
        x, w = rng.standard_normal((S, S, ci)), rng.standard_normal((co, ci, 3, 3))
        wf = np.transpose(w, (0, 2, 3, 1)).reshape(co, 9, ci)
        y = _halo_emulation(x, wf, [t % 3 - 1 for t in range(9)], [t // 3 - 1 for t in range(9)], bw, bh)
        assert np.abs(y - conv_ref(x, w)).max() < 1e-9
        dy = rng.standard_normal((S, S, co))
        wd = np.transpose(w, (1, 2, 3, 0)).reshape(ci, 9, co)
        dx = _halo_emulation(dy, wd, [1 - t % 3 for t in range(9)], [1 - t // 3 for t in range(9)], bw, bh)
        assert np.abs(dx - conv_ref(dy, np.transpose(w[:, :, ::-1, ::-1], (1, 0, 2, 3)))).max() < 1e-9
