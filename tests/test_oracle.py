"""CPU tests of the oracle (oracle/ops.c + oracle/oracle.py).

The oracle restates the reference's kernels; here it is cross-checked against an independent
implementation (PyTorch CPU fp32 ops + autograd) so that a slip in the restatement cannot hide.
The pin against the reference's own compiled kernels is tests/test_golden.py.
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import oracle as O

torch.manual_seed(0)


def t_nchw(x):  # NHWC numpy -> NCHW torch
    return torch.from_numpy(x).permute(0, 3, 1, 2).contiguous()


def n_nhwc(t):
    return t.permute(0, 2, 3, 1).contiguous().numpy()


@pytest.mark.parametrize("S,k,cin,cout,stride,N", [
    (16, 7, 3, 8, 2, 2), (8, 1, 16, 32, 1, 3), (8, 3, 16, 16, 1, 2), (8, 3, 16, 24, 2, 2), (7, 3, 8, 8, 1, 2),
    (14, 3, 8, 16, 2, 1),
])
def test_conv_ops_vs_torch(S, k, cin, cout, stride, N):
    rng = np.random.default_rng(1)
    x = rng.standard_normal((N, S, S, cin)).astype(np.float32)
    w = rng.standard_normal((cout, cin, k, k)).astype(np.float32)
    xt = t_nchw(x).requires_grad_(True)
    wt = torch.from_numpy(w).requires_grad_(True)
    yt = F.conv2d(xt, wt, stride=stride, padding=k // 2)
    y = O.conv_fwd(x, w, stride)
    assert y.shape == (N, S // stride, S // stride, cout)
    np.testing.assert_allclose(y, n_nhwc(yt.detach()), rtol=1e-4, atol=1e-4)
    dy = rng.standard_normal(y.shape).astype(np.float32)
    yt.backward(t_nchw(dy))
    dx = O.conv_dgrad(w, dy, S, stride)
    np.testing.assert_allclose(dx, n_nhwc(xt.grad), rtol=1e-4, atol=1e-4)
    base = rng.standard_normal(x.shape).astype(np.float32)
    dx2 = O.conv_dgrad(w, dy, S, stride, din=base.copy())
    np.testing.assert_allclose(dx2, base + dx, rtol=1e-5, atol=1e-5)
    dw = O.conv_wgrad(x, dy, k, stride)
    np.testing.assert_allclose(dw, wt.grad.numpy(), rtol=1e-4, atol=2e-4)


@pytest.mark.parametrize("relu", [False, True])
def test_bn_ops_vs_torch(relu):
    rng = np.random.default_rng(2)
    N, S, Cc, eps = 4, 6, 10, 1e-7
    x = (rng.standard_normal((N, S, S, Cc)) * 3 + 1.5).astype(np.float32)
    g = rng.standard_normal(Cc).astype(np.float32)
    b = rng.standard_normal(Cc).astype(np.float32)
    xt = t_nchw(x).requires_grad_(True)
    gt, bt = torch.from_numpy(g).requires_grad_(True), torch.from_numpy(b).requires_grad_(True)
    yt = F.batch_norm(xt, None, None, gt, bt, training=True, eps=eps)
    if relu:
        yt = F.relu(yt)
    mu, var, y, xh, nv = O.bn_fwd(x, g, b, eps, relu, keep=True)
    np.testing.assert_allclose(mu, x.reshape(-1, Cc).mean(0), rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(var, x.reshape(-1, Cc).var(0), rtol=1e-4, atol=1e-5)  # biased
    np.testing.assert_allclose(y, n_nhwc(yt.detach()), rtol=1e-4, atol=1e-4)
    dy = rng.standard_normal(y.shape).astype(np.float32)
    yt.backward(t_nchw(dy))
    dg, db, dx = O.bn_bwd(x, g, eps, mu, var, y, dy, relu)
    np.testing.assert_allclose(dg, gt.grad.numpy(), rtol=1e-3, atol=1e-3)
    np.testing.assert_allclose(db, bt.grad.numpy(), rtol=1e-3, atol=1e-3)
    np.testing.assert_allclose(dx, n_nhwc(xt.grad), rtol=1e-3, atol=1e-4)


def test_pool_softmax_matmul_adam():
    rng = np.random.default_rng(3)
    x = rng.standard_normal((2, 8, 8, 5)).astype(np.float32)
    y, inds = O.maxpool_fwd(x, 3, 2)
    yt = F.max_pool2d(t_nchw(x), 3, 2, 1)
    np.testing.assert_array_equal(y, n_nhwc(yt))
    assert np.array_equal(x.reshape(-1)[inds.reshape(-1)].reshape(y.shape), y)
    dy = rng.standard_normal(y.shape).astype(np.float32)
    xt = t_nchw(x).requires_grad_(True)
    F.max_pool2d(xt, 3, 2, 1).backward(t_nchw(dy))
    np.testing.assert_allclose(O.maxpool_bwd(inds, dy, x.shape), n_nhwc(xt.grad), rtol=1e-6, atol=1e-6)
    # everything below -1024 : the reference's init value wins and the index stays -1024
    lo = np.full((1, 4, 4, 1), -2000.0, np.float32)
    ylo, ilo = O.maxpool_fwd(lo, 3, 2)
    assert (ylo == -1024).all() and (ilo == -1024).all()
    # avgpool
    p = O.avgpool_fwd(x)
    np.testing.assert_allclose(p, x.mean((1, 2)), rtol=1e-5, atol=1e-6)
    dp = rng.standard_normal(p.shape).astype(np.float32)
    np.testing.assert_allclose(O.avgpool_bwd(dp, 8), np.broadcast_to(dp[:, None, None, :] / 64, x.shape), rtol=1e-6)
    # matmul in the three forms the head uses
    A = rng.standard_normal((6, 20)).astype(np.float32)
    B = rng.standard_normal((20, 9)).astype(np.float32)
    np.testing.assert_allclose(O.matmul(A, B), A @ B, rtol=1e-5, atol=1e-5)
    D = rng.standard_normal((6, 9)).astype(np.float32)
    np.testing.assert_allclose(O.matmul(A, D, ta=True), A.T @ D, rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(O.matmul(D, B, tb=True), D @ B.T, rtol=1e-5, atol=1e-5)
    # softmax / CE
    L = (rng.standard_normal((6, 9)) * 5).astype(np.float32)
    sm = O.softmax(L)
    np.testing.assert_allclose(sm, F.softmax(torch.from_numpy(L), 1).numpy(), rtol=1e-5, atol=1e-7)
    lab = rng.integers(0, 9, 6).astype(np.int32)
    d = O.ce_deriv(sm, lab)
    oh = np.eye(9, dtype=np.float32)[lab]
    np.testing.assert_allclose(d, sm - oh, atol=1e-7)
    loss, nwrong = O.loss_acc(sm, lab)
    np.testing.assert_allclose(loss, -np.log(sm[np.arange(6), lab]).sum(), rtol=1e-5)
    assert nwrong == int((sm.argmax(1) != lab).sum())
    # ties count as wrong (reference: resnet.cu:3376)
    tie = np.full((1, 4), 0.25, np.float32)
    assert O.loss_acc(tie, np.array([2], np.int32))[1] == 1
    # Adam, first two steps vs closed form
    p0 = rng.standard_normal(50).astype(np.float32)
    g = rng.standard_normal(50).astype(np.float32)
    p, m, v = p0.copy(), np.zeros(50, np.float32), np.zeros(50, np.float32)
    O.adam(p, g, m, v, 1e-3, 0.0, 0.9, 0.999, 0.9, 0.999, 1e-7)
    np.testing.assert_allclose(p, p0 - 1e-3 * g / (np.abs(g) + 1e-7), rtol=1e-5, atol=1e-7)
    gn = g.copy()
    gn[3] = np.nan
    p2, m2, v2 = p0.copy(), m.copy(), v.copy()
    O.adam(p2, gn, m2, v2, 1e-3, 0.0, 0.9, 0.999, 0.81, 0.998, 1e-7)
    assert m2[3] == m[3] and v2[3] == v[3] and np.isfinite(p2).all()  # NaN guard keeps the moments


def _torch_net(net, images, labels):
    """Independent autograd model of the reference network on the oracle's parameters."""
    P = [torch.from_numpy(p.copy()).requires_grad_(True) for p in net.params]
    eps = net.eps

    def bn(x, g, b):
        return F.batch_norm(x, None, None, g, b, training=True, eps=eps)

    x = t_nchw(images)
    x = F.relu(bn(F.conv2d(x, P[0], stride=2, padding=3), P[1], P[2]))
    x = F.max_pool2d(x, 3, 2, 1)
    li = 3
    for b in net.plan:
        r = F.relu(bn(F.conv2d(x, P[li]), P[li + 1], P[li + 2]))
        s = F.relu(bn(F.conv2d(r, P[li + 3], stride=b["stride"], padding=1), P[li + 4], P[li + 5]))
        e = bn(F.conv2d(s, P[li + 6]), P[li + 7], P[li + 8])
        if b["proj"]:
            sc = bn(F.conv2d(x, P[li + 9], stride=b["stride"], padding=b["proj_k"] // 2), P[li + 10], P[li + 11])
            li += 12
        else:
            sc = x
            li += 9
        x = F.relu(e + sc)
    pooled = x.mean((2, 3))
    logits = pooled @ P[li]
    loss = F.cross_entropy(logits, torch.from_numpy(labels).long(), reduction="sum")  # no 1/N
    loss.backward()
    return F.softmax(logits, 1).detach().numpy(), [p.grad.numpy() for p in P], loss.item()


def test_full_network_vs_autograd():
    """Forward, backward and one Adam step of a 4-block miniature (both shortcut kinds, both strides)."""
    net = O.OracleNet(32, 5, [0, 1, 0, 1, 0], batch=4, output=10, lr=1e-3)
    net.init_like_reference(seed=5)
    # break the gamma=1/beta=0 symmetry so BN parameter gradients are exercised
    rng = np.random.default_rng(6)
    for i, s in enumerate(net.shapes):
        if len(s) == 1:
            net.params[i] = (net.params[i] + 0.2 * rng.standard_normal(s)).astype(np.float32)
    img, lab = O.synthetic_batch(4, 32, seed=7, n_classes=10)
    pred = net.forward(img, lab)
    tp, tg, tloss = _torch_net(net, img, lab)
    np.testing.assert_allclose(pred, tp, rtol=2e-3, atol=1e-5)
    loss, _ = net.loss_acc()
    np.testing.assert_allclose(loss, tloss, rtol=1e-4)
    grads = net.backward()
    for i, (g, t) in enumerate(zip(grads, tg)):
        scale = max(1e-6, float(np.abs(t).max()))
        err = float(np.abs(g.reshape(t.shape) - t).max()) / scale
        assert err < 5e-3, (i, net.shapes[i], err)
    before = [p.copy() for p in net.params]
    net.update()
    assert all(np.isfinite(p).all() for p in net.params)
    assert any(np.abs(p - q).max() > 0 for p, q in zip(net.params, before))
    assert all((g == 0).all() for g in net.grads)  # reference zeroes gradients after the step
    assert abs(net.cur_b1 - 0.9) < 1e-6 and abs(net.cur_b2 - 0.999) < 1e-6


def test_param_inventory_resnet50():
    shapes = O.param_shapes(224, 16, [1 if i in (3, 7, 13) else 0 for i in range(16)])
    assert len(shapes) == 160                      # reference: 16 + 9 * n_conv_blocks
    assert sum(int(np.prod(s)) for s in shapes) == 47576128  # 45475008 conv + 53120 BN + 2048000 FC (SURVEY.md quotes 47583424: an arithmetic slip)
