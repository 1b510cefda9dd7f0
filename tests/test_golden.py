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
    grads = net.backward()
    for i, g in enumerate(grads):
        G.summary_close(gold[key + ".grads"][i], G.summary(g), 2e-3, "grad %d" % i)
    net.update()
    for i, p in enumerate(net.params):
        G.summary_close(gold[key + ".params1"][i], G.summary(p), 1e-3, "param1 %d" % i)
    net.forward(img, lab)
    net.backward()
    net.update()
    for i, p in enumerate(net.params):
        G.summary_close(gold[key + ".params2"][i], G.summary(p), 2e-3, "param2 %d" % i)
