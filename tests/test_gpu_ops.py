"""GPU parity tests, single operators, through the C-ABI (include/resnet_b200.h).

Bars: fp32 SIMT path vs oracle 1e-4 abs / 1e-4 rel (the reference's own testConvolution tolerance, reference:
resnet.cu:3109-3218); tensor-core (TF32) path vs oracle 3e-3 relative to the tensor's max magnitude (TF32 has a
10-bit mantissa; K up to 4608 accumulates in fp32); integer outputs (argmax indices) bit-exact.
"""
import os

import numpy as np
import pytest

from oracle import golden_cases as G
from oracle import oracle as O

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_b200.npz")


@pytest.fixture(scope="module")
def api():
    from resnet_b200 import api as a
    a.L()
    return a


def rel_max(a, b):
    return float(np.abs(a - b).max() / max(1e-9, np.abs(b).max()))


# ------------------------------------------------------------------------------ fp32 SIMT path == oracle (and == reference golden)
@pytest.mark.parametrize("idx", range(len(G.CONV_CASES)))
def test_conv_simt_vs_oracle(api, idx):
    S, k, cin, cout, stride, N = G.CONV_CASES[idx]
    x, w, dy, base = G.conv_inputs(idx)
    y = api.conv_forward(x, w, stride, impl=1)
    np.testing.assert_allclose(y, O.conv_fwd(x, w, stride), rtol=1e-4, atol=1e-4)
    din, dw = api.conv_backward(x, w, dy, stride, impl=1)
    np.testing.assert_allclose(din, O.conv_dgrad(w, dy, S, stride), rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(dw, O.conv_wgrad(x, dy, k, stride), rtol=1e-4, atol=2e-4)
    din2, _ = api.conv_backward(x, w, dy, stride, din_base=base, impl=1)
    np.testing.assert_allclose(din2, base + O.conv_dgrad(w, dy, S, stride), rtol=1e-4, atol=1e-4)
    _, dw_only = api.conv_backward(x, w, dy, stride, want_din=False, impl=1)  # stem: toComputeInputDeriv=false
    np.testing.assert_allclose(dw_only, dw, rtol=1e-5, atol=1e-5)
    if os.path.exists(GOLD):
        g = np.load(GOLD)
        np.testing.assert_allclose(y, g["conv%d.y" % idx], rtol=1e-4, atol=1e-4)
        np.testing.assert_allclose(din, g["conv%d.din" % idx], rtol=1e-4, atol=1e-4)
        np.testing.assert_allclose(dw, g["conv%d.dw" % idx], rtol=1e-4, atol=2e-4)


# ------------------------------------------------------------------------------ tensor-core path
TC_CASES = [  # S, k, cin, cout, stride, N  -- every (k, stride) kind of the network, ragged tiles, partial batches
    (8, 1, 64, 64, 1, 2), (8, 1, 64, 256, 1, 3), (8, 1, 256, 64, 1, 4), (8, 3, 64, 64, 1, 2), (8, 3, 128, 128, 2, 2),
    (8, 3, 256, 512, 2, 2), (14, 3, 256, 256, 1, 3), (7, 3, 512, 512, 1, 2), (7, 1, 512, 2048, 1, 4), (28, 3, 128, 128, 1, 2),
    (14, 3, 512, 512, 2, 2), (56, 3, 64, 64, 1, 1),
]


@pytest.mark.parametrize("S,k,cin,cout,stride,N", TC_CASES)
def test_conv_tensor_core_vs_oracle(api, S, k, cin, cout, stride, N):
    rng = np.random.default_rng(S * 1000 + cin + cout + k)
    x = rng.standard_normal((N, S, S, cin)).astype(np.float32)
    w = (rng.standard_normal((cout, cin, k, k)) * 0.1).astype(np.float32)
    dy = rng.standard_normal((N, S // stride, S // stride, cout)).astype(np.float32)
    base = rng.standard_normal(x.shape).astype(np.float32)
    tol = 3e-3
    y = api.conv_forward(x, w, stride, impl=0)
    assert rel_max(y, O.conv_fwd(x, w, stride)) < tol
    din, dw = api.conv_backward(x, w, dy, stride, impl=0)
    din_ref = O.conv_dgrad(w, dy, S, stride)
    assert rel_max(din, din_ref) < tol
    assert rel_max(dw, O.conv_wgrad(x, dy, k, stride)) < tol
    din2, _ = api.conv_backward(x, w, dy, stride, din_base=base, impl=0)
    assert rel_max(din2, base + din_ref) < tol


def test_conv_tensor_core_linearity_full_size(api):
    """Size-independent property at a full ResNet-50 layer shape (3x3, 14x14, 256->256, batch 32): conv(a*x1 + x2) ==
    a*conv(x1) + conv(x2), and agreement with the fp32 SIMT path on the same inputs."""
    rng = np.random.default_rng(7)
    N, S, cin, cout = 32, 14, 256, 256
    x1 = rng.standard_normal((N, S, S, cin)).astype(np.float32)
    x2 = rng.standard_normal((N, S, S, cin)).astype(np.float32)
    w = (rng.standard_normal((cout, cin, 3, 3)) * 0.05).astype(np.float32)
    y1, y2 = api.conv_forward(x1, w, 1, impl=0), api.conv_forward(x2, w, 1, impl=0)
    y12 = api.conv_forward((2.0 * x1 + x2).astype(np.float32), w, 1, impl=0)
    assert rel_max(y12, 2.0 * y1 + y2) < 4e-3
    assert rel_max(y1, api.conv_forward(x1, w, 1, impl=1)) < 3e-3


@pytest.mark.parametrize("S,N", [(32, 4), (64, 3), (224, 2)])
def test_stem_tensor_core_vs_oracle(api, S, N):
    """7x7/2, Cin = 3 stem on the tcgen05 path (zero-bordered NHWC4 copy + overlapping-row tensor maps): fprop and wgrad."""
    rng = np.random.default_rng(S + N)
    x, _ = O.synthetic_batch(N, S, seed=S)
    w = rng.normal(0, np.sqrt(2.0 / (49 * 67)), (64, 3, 7, 7)).astype(np.float32)
    dy = rng.standard_normal((N, S // 2, S // 2, 64)).astype(np.float32)
    y = api.conv_forward(x, w, 2, impl=0)
    assert rel_max(y, O.conv_fwd(x, w, 2)) < 3e-3
    _, dw = api.conv_backward(x, w, dy, 2, want_din=False, impl=0)
    assert rel_max(dw, O.conv_wgrad(x, dy, 7, 2)) < 3e-3


# ------------------------------------------------------------------------------ bandwidth-bound kernels
@pytest.mark.parametrize("idx", range(len(G.BN_CASES)))
def test_batchnorm_vs_oracle(api, idx):
    x, g, b, dy, relu = G.bn_inputs(idx)
    mu, var, y = api.batchnorm_forward(x, g, b, 1e-7, relu)
    omu, ovar, oy, _, _ = O.bn_fwd(x, g, b, 1e-7, relu)
    np.testing.assert_allclose(mu, omu, rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(var, ovar, rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(y, oy, rtol=1e-4, atol=1e-4)
    dg, db, dx = api.batchnorm_backward(x, g, 1e-7, omu, ovar, oy, dy, relu)
    odg, odb, odx = O.bn_bwd(x, g, 1e-7, omu, ovar, oy, dy, relu)
    np.testing.assert_allclose(dg, odg, rtol=1e-4, atol=1e-3)
    np.testing.assert_allclose(db, odb, rtol=1e-4, atol=1e-3)
    np.testing.assert_allclose(dx, odx, rtol=1e-3, atol=1e-4)
    if os.path.exists(GOLD):
        gold = np.load(GOLD)
        np.testing.assert_allclose(y, gold["bn%d.activated" % idx], rtol=1e-4, atol=1e-4)
        np.testing.assert_allclose(dx, gold["bn%d.dx" % idx], rtol=1e-3, atol=1e-4)


@pytest.mark.parametrize("N,S,Cc,relu", [(4, 8, 128, 1), (8, 16, 64, 1), (4, 8, 512, 0), (16, 14, 256, 1), (32, 28, 128, 1), (3, 7, 2048, 1)])
def test_batchnorm_multi_block_geometries(api, N, S, Cc, relu):
    """geometries whose streams span several thread blocks per column group (the in-block and cross-block folds)"""
    rng = np.random.default_rng(N * 1000 + S * 10 + Cc)
    x = (rng.standard_normal((N, S, S, Cc)) * 1.5 + 0.3).astype(np.float32)
    g = (1 + 0.2 * rng.standard_normal(Cc)).astype(np.float32)
    b = (0.2 * rng.standard_normal(Cc)).astype(np.float32)
    dy = rng.standard_normal(x.shape).astype(np.float32)
    mu, var, y = api.batchnorm_forward(x, g, b, 1e-7, relu)
    omu, ovar, oy, _, _ = O.bn_fwd(x, g, b, 1e-7, relu)
    np.testing.assert_allclose(mu, omu, rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(var, ovar, rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(y, oy, rtol=1e-4, atol=1e-4)
    dg, db, dx = api.batchnorm_backward(x, g, 1e-7, omu, ovar, oy, dy, relu)
    odg, odb, odx = O.bn_bwd(x, g, 1e-7, omu, ovar, oy, dy, relu)
    np.testing.assert_allclose(db, odb, rtol=1e-4, atol=1e-3)
    np.testing.assert_allclose(dg, odg, rtol=1e-4, atol=1e-3)
    np.testing.assert_allclose(dx, odx, rtol=1e-3, atol=1e-4)


def test_batchnorm_wide_channels_and_residual(api):
    """C = 2048 (two column groups per 256-thread block) and the fused residual + ReLU join."""
    rng = np.random.default_rng(11)
    x = rng.standard_normal((3, 7, 7, 2048)).astype(np.float32)
    res = rng.standard_normal(x.shape).astype(np.float32)
    g = (1 + 0.1 * rng.standard_normal(2048)).astype(np.float32)
    b = (0.1 * rng.standard_normal(2048)).astype(np.float32)
    mu, var, y = api.batchnorm_forward(x, g, b, 1e-7, True, residual=res)
    omu, ovar, on, _, _ = O.bn_fwd(x, g, b, 1e-7, False)
    oy, _ = O.add_relu(on, res)
    np.testing.assert_allclose(var, ovar, rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(y, oy, rtol=1e-4, atol=1e-4)
    dy = rng.standard_normal(x.shape).astype(np.float32)
    dg, db, dx = api.batchnorm_backward(x, g, 1e-7, omu, ovar, oy, dy, True)
    odg, odb, odx = O.bn_bwd(x, g, 1e-7, omu, ovar, oy, dy, True)
    np.testing.assert_allclose(dg, odg, rtol=1e-3, atol=1e-3)
    np.testing.assert_allclose(dx, odx, rtol=1e-3, atol=1e-4)


def test_tf32_rounding_is_idempotent(api):
    rng = np.random.default_rng(12)
    x = rng.standard_normal((2, 4, 4, 64)).astype(np.float32)
    g, b = np.ones(64, np.float32), np.zeros(64, np.float32)
    _, _, y = api.batchnorm_forward(x, g, b, 1e-7, False, round_tf32=True)
    assert (y.view(np.uint32) & 0x1FFF == 0).all()          # 13 low mantissa bits cleared
    _, _, y0 = api.batchnorm_forward(x, g, b, 1e-7, False)
    assert np.abs(y - y0).max() <= np.abs(y0).max() * 2.0 ** -11


def test_maxpool_bit_exact(api):
    x = G.maxpool_input()
    out, inds = api.maxpool_forward(x, 3, 2)
    oo, oi = O.maxpool_fwd(x, 3, 2)
    np.testing.assert_array_equal(out, oo)
    np.testing.assert_array_equal(inds, oi)                  # first max in row-major scan wins (ties seeded in the input)
    if os.path.exists(GOLD):
        gold = np.load(GOLD)
        np.testing.assert_array_equal(inds, gold["maxpool.inds"])
        np.testing.assert_array_equal(out, gold["maxpool.out"])
    rng = np.random.default_rng(5)
    x64 = rng.standard_normal((3, 16, 16, 64)).astype(np.float32)
    out, inds = api.maxpool_forward(x64, 3, 2)
    oo, oi = O.maxpool_fwd(x64, 3, 2)
    np.testing.assert_array_equal(inds, oi)
    dy = rng.standard_normal(out.shape).astype(np.float32)
    np.testing.assert_allclose(api.maxpool_backward(inds, dy, x64.shape, 3, 2), O.maxpool_bwd(oi, dy, x64.shape), rtol=1e-6, atol=1e-6)


def test_head_ops(api):
    rng = np.random.default_rng(6)
    x = rng.standard_normal((5, 7, 7, 128)).astype(np.float32)
    np.testing.assert_allclose(api.avgpool_forward(x), O.avgpool_fwd(x), rtol=1e-5, atol=1e-6)
    dp = rng.standard_normal((5, 128)).astype(np.float32)
    np.testing.assert_allclose(api.avgpool_backward(dp, 7), O.avgpool_bwd(dp, 7), rtol=1e-6)
    A = rng.standard_normal((32, 2048)).astype(np.float32)
    B = (rng.standard_normal((2048, 1000)) * 0.01).astype(np.float32)
    np.testing.assert_allclose(api.matmul(A, B), O.matmul(A, B), rtol=1e-4, atol=1e-5)   # reference testMatMul shape, tol 1e-5
    D = rng.standard_normal((32, 1000)).astype(np.float32)
    np.testing.assert_allclose(api.matmul(A, D, ta=True), O.matmul(A, D, ta=True), rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(api.matmul(D, B, tb=True), O.matmul(D, B, tb=True), rtol=1e-4, atol=1e-5)
    logits = (rng.standard_normal((9, 1000)) * 4).astype(np.float32)
    labels = rng.integers(0, 1000, 9).astype(np.int32)
    pred, d = api.softmax_ce(logits, labels)
    op = O.softmax(logits)
    np.testing.assert_allclose(pred, op, rtol=1e-5, atol=1e-8)
    np.testing.assert_allclose(d, O.ce_deriv(op, labels), rtol=1e-5, atol=1e-7)
    assert (pred.argmax(1) == op.argmax(1)).all()


def test_adam_vs_oracle_and_reference(api):
    p, g1, g2 = G.adam_inputs()
    m, v = np.zeros_like(p), np.zeros_like(p)
    op, om, ov = p.copy(), m.copy(), v.copy()
    p1, gz, m1, v1 = api.adam(p, g1, m, v, 1e-3, 0.0, 0.9, 0.999, 0.9, 0.999, 1e-7)
    O.adam(op, g1, om, ov, 1e-3, 0.0, 0.9, 0.999, 0.9, 0.999, 1e-7)
    np.testing.assert_allclose(p1, op, rtol=1e-6, atol=1e-7)
    assert (gz == 0).all()                                   # gradients are zeroed by the step (reference: resnet.cu:2972-2975)
    p2, _, m2, v2 = api.adam(p1, g2, m1, v1, 1e-3, 0.01, 0.9, 0.999, 0.81, 0.998001, 1e-7)
    O.adam(op, g2, om, ov, 1e-3, 0.01, 0.9, 0.999, 0.81, 0.998001, 1e-7)
    np.testing.assert_allclose(p2, op, rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(m2, om, rtol=1e-6, atol=1e-8)
    assert np.isfinite(p2).all() and m2[5] == m1[5] and v2[9] == v1[9]   # NaN / Inf gradients keep the moments
    if os.path.exists(GOLD):
        gold = np.load(GOLD)
        np.testing.assert_allclose(p1, gold["adam.p1"], rtol=1e-6, atol=1e-7)
        np.testing.assert_allclose(p2, gold["adam.p2"], rtol=1e-5, atol=1e-7)
