"""Pins the oracle against the reference's OWN compiled kernels.

tests/golden/reference_b200.npz was produced on a B200 by oracle/gen_golden.py, which drives the reference's
translation units (resnet.cu, resnet_clean.cu, resnet_cudnn.cu compiled from /root/reference by oracle/Makefile)
on the deterministic cases of oracle/golden_cases.py.  Here the oracle replays the same cases on the CPU.

Tolerances: the reference's own self-test bars (1e-4 abs for conv, 1e-5 for GEMM-like ops; reference:
resnet.cu:3033-3218); whole-network fingerprints 1e-3 (fp32 re-association across ~10 layers).
"""
import os

import numpy as np
import pytest

from oracle import golden_cases as G
from oracle import oracle as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_b200.npz")
pytestmark = pytest.mark.skipif(not os.path.exists(GOLD), reason="golden fixture not generated yet")


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


@pytest.mark.parametrize("idx", range(len(G.CONV_CASES)))
def test_conv_kernels(gold, idx):
    S, k, cin, cout, stride, N = G.CONV_CASES[idx]
    x, w, dy, base = G.conv_inputs(idx)
    np.testing.assert_allclose(O.conv_fwd(x, w, stride), gold["conv%d.y" % idx], rtol=1e-5, atol=1e-4)
    np.testing.assert_allclose(O.conv_dgrad(w, dy, S, stride), gold["conv%d.din" % idx], rtol=1e-5, atol=1e-4)
    np.testing.assert_allclose(O.conv_dgrad(w, dy, S, stride, din=base.copy()), gold["conv%d.din_add" % idx], rtol=1e-5, atol=1e-4)
    np.testing.assert_allclose(O.conv_wgrad(x, dy, k, stride), gold["conv%d.dw" % idx], rtol=1e-5, atol=1e-4)


@pytest.mark.parametrize("idx", range(len(G.BN_CASES)))
def test_batchnorm_kernels(gold, idx):
    x, g, b, dy, relu = G.bn_inputs(idx)
    mu, var, act, xh, nv = O.bn_fwd(x, g, b, 1e-7, relu, keep=True)
    for nm, v in (("means", mu), ("vars", var), ("xhat", xh), ("normalized", nv), ("activated", act)):
        np.testing.assert_allclose(v, gold["bn%d.%s" % (idx, nm)], rtol=1e-5, atol=1e-5, err_msg=nm)
    dg, db, dx = O.bn_bwd(x, g, 1e-7, mu, var, act, dy, relu)
    np.testing.assert_allclose(dg, gold["bn%d.dgamma" % idx], rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(db, gold["bn%d.dbeta" % idx], rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(dx, gold["bn%d.dx" % idx], rtol=1e-4, atol=1e-5)


def test_maxpool_and_adam_kernels(gold):
    out, inds = O.maxpool_fwd(G.maxpool_input(), 3, 2)
    np.testing.assert_array_equal(out, gold["maxpool.out"])
    np.testing.assert_array_equal(inds, gold["maxpool.inds"])
    p, g1, g2 = G.adam_inputs()
    m, v = np.zeros_like(p), np.zeros_like(p)
    O.adam(p, g1, m, v, 1e-3, 0.0, 0.9, 0.999, 0.9, 0.999, 1e-7)
    np.testing.assert_allclose(p, gold["adam.p1"], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(m, gold["adam.m1"], rtol=1e-6, atol=1e-8)
    O.adam(p, g2, m, v, 1e-3, 0.01, 0.9, 0.999, 0.81, 0.998001, 1e-7)
    np.testing.assert_allclose(p, gold["adam.p2"], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(m, gold["adam.m2"], rtol=1e-6, atol=1e-8)
    np.testing.assert_allclose(v, gold["adam.v2"], rtol=1e-6, atol=1e-8)


@pytest.mark.parametrize("tag", ["mini", "mini4", "mini5"])
def test_network_forward_vs_resnet_cu(gold, tag):
    """forward_pass of the reference's resnet.cu, layer by layer (fingerprints) and pred (full)."""
    cfg = {"mini": G.MINI, "mini4": G.MINI4, "mini5": G.MINI5}[tag]
    net = O.OracleNet(cfg["input_dim"], cfg["n_blocks"], cfg["reductions"], cfg["batch"], output=cfg["output"], lr=cfg["lr"])
    shapes = O.param_shapes(cfg["input_dim"], cfg["n_blocks"], cfg["reductions"], output=cfg["output"])
    net.set_params(G.mini_weights(shapes))
    img, lab = G.mini_batch(cfg)
    pred = net.forward(img, lab)
    key = "%s.naive" % tag
    if key + ".pred" not in gold.files:
        pytest.skip("network stage missing from the fixture")
    np.testing.assert_allclose(pred, gold[key + ".pred"], rtol=1e-3, atol=1e-6)
    assert (pred.argmax(1) == gold[key + ".pred"].argmax(1)).all()
    np.testing.assert_array_equal(net.act["max_inds"].reshape(-1), gold[key + ".max_inds"])
    checked = 0
    for k in gold.files:
        if k.startswith(key + ".act."):
            nm = k[len(key + ".act."):]
            G.summary_close(gold[k], G.summary(net.act[nm]), 1e-3, nm)
            checked += 1
    assert checked > 20


@pytest.mark.parametrize("tag,variant", [("mini5", "clean"), ("mini5", "cudnn")])
def test_network_step_vs_complete_variants(gold, tag, variant):
    """full step (forward, backward, Adam x2) of resnet_clean.cu / resnet_cudnn.cu: pred, every parameter gradient,
    parameters after one and two updates."""
    key = "%s.%s" % (tag, variant)
    if key + ".grads" not in gold.files:
        pytest.skip("variant did not run on the generating box")
    cfg = G.MINI5
    net = O.OracleNet(cfg["input_dim"], cfg["n_blocks"], cfg["reductions"], cfg["batch"], output=cfg["output"], lr=cfg["lr"])
    shapes = O.param_shapes(cfg["input_dim"], cfg["n_blocks"], cfg["reductions"], output=cfg["output"])
    net.set_params(G.mini_weights(shapes))
    img, lab = G.mini_batch(cfg)
    pred = net.forward(img, lab)
    np.testing.assert_allclose(pred, gold[key + ".pred"], rtol=2e-3, atol=1e-6)
    assert (pred.argmax(1) == gold[key + ".pred"].argmax(1)).all()
    grads = net.backward()
    ref = gold[key + ".grads"]
    nloc = len(grads)

    def l2_close(i, tol):
        r, o = np.sqrt(ref[i][1]), np.sqrt(G.summary(grads[i])[1])
        assert abs(r - o) <= tol * max(r, 1e-12), ("grad l2", i, shapes[i], r, o)

    if variant == "cudnn":
        # What this variant can and cannot pin, as observed on the generating B200 (cuDNN 9.10):
        #  * cuDNN's default math mode runs fp32 convolutions on TF32 tensor cores, so element values carry TF32 noise
        #    (and the ReLU-mask flips it causes); per-tensor L2 norms still agree to ~1e-4.
        #  * locations 0-2 (stem conv / BN) sit below maxPoolDeriv, which OVERWRITES where 3x3/2 windows overlap
        #    (reference: resnet.cu:493); we accumulate (as cudnnPoolingBackward does), so they differ.
        for i in range(3, nloc):
            l2_close(i, 5e-3)
            head_r, head_o = ref[i][2:], G.summary(grads[i])[2:]
            assert np.abs(head_r - head_o).max() <= 0.1 * max(np.abs(head_r).max(), 1e-9), ("grad head", i)
        net.update()
        p1 = gold[key + ".params1"]
        for i in range(3, nloc):
            r, o = np.sqrt(p1[i][1]), np.sqrt(G.summary(net.params[i])[1])
            assert abs(r - o) <= 2e-3 * max(r, 1e-12), ("param l2 after Adam", i)
    else:
        # resnet_clean.cu: forward agrees to 1e-7 and so does the FC gradient, but below the head its gradient norms grow
        # by ~30x per layer (1e6 at the stem vs 2e3 from autograd / the cuDNN variant / us): its backward is broken at these
        # sizes, so it cannot serve as a backward oracle.  Recorded here so the discrepancy stays visible.
        l2_close(nloc - 1, 1e-4)
        assert np.sqrt(ref[0][1]) > 100 * np.sqrt(G.summary(grads[0])[1])
