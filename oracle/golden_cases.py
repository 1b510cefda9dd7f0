"""oracle/golden_cases.py -- TEST INFRASTRUCTURE.  Deterministic inputs shared by the generator
(oracle/gen_golden.py, run on a GPU box against the reference's compiled kernels) and by the tests that
replay the same cases through the oracle and through libresnet_b200.so.  Inputs are re-derived from
seeds, so the committed fixture (tests/golden/reference_b200.npz) only holds the reference's OUTPUTS."""
import numpy as np

# (S, k, cin, cout, stride, N) -- covers 7x7/2 stem-like, 1x1, 3x3/1, 3x3/2, odd spatial (7), Cin=3
CONV_CASES = [(16, 7, 3, 8, 2, 2), (8, 1, 16, 32, 1, 4), (8, 3, 16, 16, 1, 2), (8, 3, 16, 24, 2, 2), (14, 3, 8, 16, 2, 1),
              (7, 3, 8, 8, 1, 2)]
BN_CASES = [(4, 6, 10, 0), (4, 6, 10, 1), (2, 4, 64, 1)]  # (N, S, C, relu)
# NB: the last block of every network config is non-strided: the reference pools the final activation over the LAST
# BLOCK'S *incoming* spatial dim (reference: resnet.cu:1732), which equals the output dim only then (true for ResNet-50).
MINI = dict(input_dim=32, n_blocks=3, reductions=[0, 1, 0], batch=4, output=10, lr=1e-3, wd=0.0, b1=0.9, b2=0.999, eps=1e-7)
# MINI5 has exactly four projection blocks, so the reference's n_locations = 16 + 9*n_blocks (reference: resnet.cu:819) is exact
# and its update_parameters / check_errors loops (resnet.cu:2952) do not run off the end of locations[].
MINI5 = dict(input_dim=32, n_blocks=5, reductions=[0, 1, 1, 1, 0], batch=4, output=10, lr=1e-3, wd=0.0, b1=0.9, b2=0.999, eps=1e-7)
MINI4 = dict(input_dim=32, n_blocks=4, reductions=[0, 1, 0, 0], batch=2, output=10, lr=1e-3, wd=0.0, b1=0.9, b2=0.999, eps=1e-7)


def conv_inputs(idx):
    S, k, cin, cout, stride, N = CONV_CASES[idx]
    rng = np.random.default_rng(100 + idx)
    x = rng.standard_normal((N, S, S, cin)).astype(np.float32)
    w = (rng.standard_normal((cout, cin, k, k)) * 0.2).astype(np.float32)
    dy = rng.standard_normal((N, S // stride, S // stride, cout)).astype(np.float32)
    base = rng.standard_normal(x.shape).astype(np.float32)
    return x, w, dy, base


def bn_inputs(idx):
    N, S, Cc, relu = BN_CASES[idx]
    rng = np.random.default_rng(200 + idx)
    x = (rng.standard_normal((N, S, S, Cc)) * 2 + 0.7).astype(np.float32)
    g = (1 + 0.3 * rng.standard_normal(Cc)).astype(np.float32)
    b = (0.3 * rng.standard_normal(Cc)).astype(np.float32)
    dy = rng.standard_normal(x.shape).astype(np.float32)
    return x, g, b, dy, relu


def maxpool_input():
    rng = np.random.default_rng(300)
    x = rng.standard_normal((2, 8, 8, 5)).astype(np.float32)
    x[0, 0:3, 0:3, 0] = 0.5  # ties: first max in row-major window scan must win
    return x


def adam_inputs():
    rng = np.random.default_rng(400)
    p = rng.standard_normal(64).astype(np.float32)
    g1 = rng.standard_normal(64).astype(np.float32)
    g2 = rng.standard_normal(64).astype(np.float32)
    g2[5] = np.nan
    g2[9] = np.inf
    return p, g1, g2


def mini_weights(shapes, seed=11):
    """numpy-seeded weights in locations[] order (reference init distribution, plus perturbed gamma/beta so
    BatchNorm parameter gradients are exercised)."""
    rng = np.random.default_rng(seed)
    out = []
    for i, s in enumerate(shapes):
        if len(s) == 4:
            out.append(rng.normal(0, np.sqrt(2.0 / (s[2] * s[3] * (s[0] + s[1]))), s).astype(np.float32))
        elif len(s) == 2:
            out.append(rng.normal(0, 1e-2, s).astype(np.float32))
        else:
            is_gamma = len(shapes[i - 1]) == 4
            out.append(((1.0 if is_gamma else 0.0) + 0.2 * rng.standard_normal(s)).astype(np.float32))
    return out


def mini_batch(cfg, seed=21):
    rng = np.random.default_rng(seed)
    img = rng.integers(0, 256, size=(cfg["batch"], cfg["input_dim"], cfg["input_dim"], 3)).astype(np.float32)
    img -= np.array([103.94, 116.78, 123.68], np.float32)
    lab = rng.integers(0, cfg["output"], size=cfg["batch"]).astype(np.int32)
    return np.ascontiguousarray(img), lab


def summary(a):
    """order-insensitive + positional fingerprint of a tensor: [sum, sum of squares, first 16 values]"""
    a = np.asarray(a, np.float64).reshape(-1)
    head = np.zeros(16)
    head[:min(16, a.size)] = a[:16]
    return np.concatenate([[a.sum(), (a * a).sum()], head])


def summary_close(s_ref, s_got, rtol, name=""):
    scale = np.sqrt(max(s_ref[1], 1e-30))  # l2 norm of the tensor
    assert abs(s_ref[0] - s_got[0]) <= rtol * max(scale, abs(s_ref[0])) * 4 + 1e-6, (name, "sum", s_ref[0], s_got[0])
    assert abs(np.sqrt(s_ref[1]) - np.sqrt(max(s_got[1], 0))) <= rtol * scale + 1e-6, (name, "l2", s_ref[1], s_got[1])
    head_scale = max(np.abs(s_ref[2:]).max(), 1e-6)
    assert np.abs(s_ref[2:] - s_got[2:]).max() <= rtol * head_scale * 4 + 1e-6, (name, "head", s_ref[2:6], s_got[2:6])


ACT_NAMES_FWD = ["init_conv_applied", "norm_init_conv.means", "norm_init_conv.vars", "init_conv_activated",
                 "init_convblock_input", "final_conv_output_pooled", "linear_output"]
BLOCK_FIELDS_FWD = ["post_reduced", "norm_post_reduced.means", "norm_post_reduced.vars", "post_reduced_activated",
                    "post_spatial", "norm_post_spatial.means", "norm_post_spatial.vars", "post_spatial_activated",
                    "post_expanded", "norm_post_expanded.means", "norm_post_expanded.vars", "post_expanded_norm_vals",
                    "transformed_residual", "output", "output_activated"]
