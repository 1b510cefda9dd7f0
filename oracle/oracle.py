"""oracle/oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

ctypes bindings for oracle/ops.c plus the network wiring of the reference's training step,
restated on the host: forward (reference: resnet.cu:1526-1775), backward (reference:
resnet.cu:1777-2248, with the spatial BatchNorm backward that resnet.cu:2060-2083 forgets to
launch, as done by resnet_clean.cu:2778 / resnet_cudnn.cu:2365) and Adam update (reference:
resnet.cu:2910-2987).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
reference legs may import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

f32p = C.POINTER(C.c_float)
i32p = C.POINTER(C.c_int)


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "ops.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.oracle_num_threads.restype = C.c_int
    return _LIB


def _p(a):
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"]
    if a.dtype == np.float32:
        return a.ctypes.data_as(f32p)
    if a.dtype == np.int32:
        return a.ctypes.data_as(i32p)
    raise TypeError(a.dtype)


def num_threads():
    return lib().oracle_num_threads()


# ----------------------------------------------------------------------------- per-op wrappers
def conv_fwd(x, w, stride):
    """x: [N,S,S,Cin] NHWC, w: [Cout,Cin,k,k] -> [N,S/stride,S/stride,Cout]"""
    N, S, _, cin = x.shape
    cout, _, k, _ = w.shape
    out = np.empty((N, S // stride, S // stride, cout), np.float32)
    lib().oracle_conv_fwd(_p(x), _p(w), S, k, cin, cout, stride, N, _p(out))
    return out


def conv_dgrad(w, dout, S, stride, din=None):
    cout, cin, k, _ = w.shape
    N = dout.shape[0]
    to_add = din is not None
    if din is None:
        din = np.empty((N, S, S, cin), np.float32)
    lib().oracle_conv_dgrad(_p(w), _p(dout), S, k, cin, cout, stride, N, int(to_add), _p(din))
    return din


def conv_wgrad(x, dout, k, stride):
    N, S, _, cin = x.shape
    cout = dout.shape[3]
    dw = np.empty((cout, cin, k, k), np.float32)
    lib().oracle_conv_wgrad(_p(x), _p(dout), S, k, cin, cout, stride, N, _p(dw))
    return dw


def bn_fwd(x, gamma, beta, eps, relu, keep=False):
    N, S, _, Cc = x.shape
    means = np.empty(Cc, np.float32)
    vars_ = np.empty(Cc, np.float32)
    act = np.empty_like(x)
    xhat = np.empty_like(x) if keep else None
    norm = np.empty_like(x) if keep else None
    lib().oracle_bn_fwd(_p(x), _p(gamma), _p(beta), S, Cc, N, C.c_float(eps), _p(means), _p(vars_), _p(xhat), _p(norm),
                        _p(act), int(relu))
    return means, vars_, act, xhat, norm


def bn_bwd(x, gamma, eps, means, vars_, mask_src, dy, relu):
    N, S, _, Cc = x.shape
    dgamma = np.empty(Cc, np.float32)
    dbeta = np.empty(Cc, np.float32)
    dx = np.empty_like(x)
    lib().oracle_bn_bwd(_p(x), _p(gamma), S, Cc, N, C.c_float(eps), _p(means), _p(vars_), _p(mask_src), _p(dy),
                        _p(dgamma), _p(dbeta), _p(dx), int(relu))
    return dgamma, dbeta, dx


def maxpool_fwd(x, k, stride):
    N, S, _, Cc = x.shape
    So = S // stride
    out = np.empty((N, So, So, Cc), np.float32)
    inds = np.empty((N, So, So, Cc), np.int32)
    lib().oracle_maxpool_fwd(_p(x), k, stride, S, Cc, N, _p(inds), _p(out))
    return out, inds


def maxpool_bwd(inds, dout, in_shape):
    din = np.zeros(in_shape, np.float32)
    lib().oracle_maxpool_bwd(_p(inds), _p(dout), C.c_size_t(dout.size), _p(din))
    return din


def avgpool_fwd(x):
    N, S, _, Cc = x.shape
    out = np.empty((N, Cc), np.float32)
    lib().oracle_avgpool_fwd(_p(x), S, Cc, N, _p(out))
    return out


def avgpool_bwd(dp, S):
    N, Cc = dp.shape
    din = np.empty((N, S, S, Cc), np.float32)
    lib().oracle_avgpool_bwd(_p(dp), S, Cc, N, _p(din))
    return din


def matmul(A, B, ta=False, tb=False):
    """out[m,n] = op(A)[m,k] . op(B)[k,n];  ta: A stored [k,m];  tb: B stored [n,k]"""
    m, k = (A.shape[1], A.shape[0]) if ta else A.shape
    n = B.shape[0] if tb else B.shape[1]
    out = np.empty((m, n), np.float32)
    lib().oracle_matmul(_p(A), _p(B), m, k, n, int(ta), int(tb), _p(out))
    return out


def softmax(X):
    out = np.empty_like(X)
    lib().oracle_softmax(_p(X), X.shape[0], X.shape[1], _p(out))
    return out


def ce_deriv(pred, labels):
    d = np.empty_like(pred)
    lib().oracle_ce_deriv(_p(pred), _p(labels), pred.shape[0], pred.shape[1], _p(d))
    return d


def loss_acc(pred, labels):
    ls = C.c_float()
    nw = C.c_int()
    lib().oracle_loss_acc(_p(pred), _p(labels), pred.shape[0], pred.shape[1], C.byref(ls), C.byref(nw))
    return ls.value, nw.value


def adam(p, g, m, v, lr, wd, b1, b2, cur_b1, cur_b2, eps):
    lib().oracle_adam(_p(p), _p(g), _p(m), _p(v), C.c_size_t(p.size), C.c_float(lr), C.c_float(wd), C.c_float(b1),
                      C.c_float(b2), C.c_float(cur_b1), C.c_float(cur_b2), C.c_float(eps))


def add_relu(a, b, keep_sum=False):
    act = np.empty_like(a)
    s = np.empty_like(a) if keep_sum else None
    lib().oracle_add_relu(_p(a), _p(b), C.c_size_t(a.size), _p(s), _p(act))
    return act, s


def relu_bwd(pre, up):
    out = np.empty_like(up)
    lib().oracle_relu_bwd(_p(pre), _p(up), C.c_size_t(up.size), _p(out))
    return out


# ----------------------------------------------------------------------------- network wiring
def block_plan(input_dim, n_blocks, reductions, init_filters=64):
    """Per-block shapes, reference: resnet.cu:857-927 (stride-2 blocks halve the spatial dim at
    the 3x3 and use a 3x3/2 projection; channel-change at stride 1 uses a 1x1 projection)."""
    plan = []
    incoming, spatial = init_filters, input_dim // 4
    reduced, expanded = init_filters, 4 * init_filters
    for i in range(n_blocks):
        stride = 1
        if reductions[i]:
            stride, reduced, expanded = 2, reduced * 2, expanded * 2
        plan.append(dict(incoming=incoming, spatial=spatial, reduced=reduced, expanded=expanded, stride=stride,
                         proj=(incoming != expanded), proj_k=(3 if stride == 2 else 1)))
        if reductions[i]:
            spatial //= 2
        incoming = expanded
    return plan


def param_shapes(input_dim, n_blocks, reductions, init_filters=64, init_k=7, output=1000):
    """Shapes in `locations[]` order, reference: resnet.cu:839-943."""
    shapes = [(init_filters, 3, init_k, init_k), (init_filters,), (init_filters,)]
    plan = block_plan(input_dim, n_blocks, reductions, init_filters)
    for b in plan:
        shapes += [(b["reduced"], b["incoming"], 1, 1), (b["reduced"],), (b["reduced"],)]
        shapes += [(b["reduced"], b["reduced"], 3, 3), (b["reduced"],), (b["reduced"],)]
        shapes += [(b["expanded"], b["reduced"], 1, 1), (b["expanded"],), (b["expanded"],)]
        if b["proj"]:
            shapes += [(b["expanded"], b["incoming"], b["proj_k"], b["proj_k"]), (b["expanded"],), (b["expanded"],)]
    shapes.append((plan[-1]["expanded"], output))
    return shapes


class OracleNet:
    """Host restatement of the reference trainer. params/grads/m/v are lists in locations[] order."""

    def __init__(self, input_dim, n_blocks, reductions, batch, init_filters=64, output=1000, lr=1e-4, wd=0.0,
                 b1=0.9, b2=0.999, eps=1e-7, pool_k=3, pool_stride=2, init_k=7, init_stride=2):
        self.input_dim, self.n_blocks, self.reductions = input_dim, n_blocks, list(reductions)
        self.batch, self.init_filters, self.output = batch, init_filters, output
        self.lr, self.wd, self.b1, self.b2, self.eps = lr, wd, b1, b2, eps
        self.pool_k, self.pool_stride, self.init_k, self.init_stride = pool_k, pool_stride, init_k, init_stride
        self.cur_b1 = self.cur_b2 = 1.0
        self.plan = block_plan(input_dim, n_blocks, reductions, init_filters)
        # reference: resnet.cu:1732 pools over the last block's INCOMING spatial dim; only a non-strided last block is well-defined
        assert self.plan[-1]["stride"] == 1, "last block must not be a spatial-reduction block"
        self.shapes = param_shapes(input_dim, n_blocks, reductions, init_filters, init_k, output)
        self.params = [np.zeros(s, np.float32) for s in self.shapes]
        self.m = [np.zeros(s, np.float32) for s in self.shapes]
        self.v = [np.zeros(s, np.float32) for s in self.shapes]
        self.grads = [np.zeros(s, np.float32) for s in self.shapes]
        self.act = {}
        self.dact = {}

    def set_params(self, arrays):
        assert len(arrays) == len(self.shapes)
        self.params = [np.ascontiguousarray(a, np.float32).reshape(s) for a, s in zip(arrays, self.shapes)]

    def init_like_reference(self, seed=0):
        """N(0, 2/(fan_in+fan_out)) weights, FC N(0,1e-4), gamma 1, beta 0 (reference: resnet.cu:730-938).
        (numpy draws, not cuRAND's XORWOW -- for oracle-only tests.)"""
        rng = np.random.default_rng(seed)
        for i, s in enumerate(self.shapes):
            if len(s) == 4:
                fan = s[2] * s[3] * (s[0] + s[1])
                self.params[i] = rng.normal(0, np.sqrt(2.0 / fan), s).astype(np.float32)
            elif len(s) == 2:
                self.params[i] = rng.normal(0, np.sqrt(1e-4), s).astype(np.float32)
            else:
                # after each conv weight: gamma then beta
                self.params[i] = (np.ones(s) if len(self.shapes[i - 1]) == 4 else np.zeros(s)).astype(np.float32)

    # -- forward, reference: resnet.cu:1526-1775
    def forward(self, images, labels):
        a = self.act = {"images": images, "labels": labels}
        P, eps = self.params, self.eps
        a["init_conv_applied"] = conv_fwd(images, P[0], self.init_stride)
        mu, var, y, _, _ = bn_fwd(a["init_conv_applied"], P[1], P[2], eps, True)
        a["norm_init_conv.means"], a["norm_init_conv.vars"], a["init_conv_activated"] = mu, var, y
        a["init_convblock_input"], a["max_inds"] = maxpool_fwd(y, self.pool_k, self.pool_stride)
        x = a["init_convblock_input"]
        li = 3
        for i, b in enumerate(self.plan):
            pre = "b%d." % i
            a[pre + "post_reduced"] = conv_fwd(x, P[li], 1)
            mu, var, y, _, _ = bn_fwd(a[pre + "post_reduced"], P[li + 1], P[li + 2], eps, True)
            a[pre + "norm_post_reduced.means"], a[pre + "norm_post_reduced.vars"], a[pre + "post_reduced_activated"] = mu, var, y
            a[pre + "post_spatial"] = conv_fwd(y, P[li + 3], b["stride"])
            mu, var, y, _, _ = bn_fwd(a[pre + "post_spatial"], P[li + 4], P[li + 5], eps, True)
            a[pre + "norm_post_spatial.means"], a[pre + "norm_post_spatial.vars"], a[pre + "post_spatial_activated"] = mu, var, y
            a[pre + "post_expanded"] = conv_fwd(y, P[li + 6], 1)
            mu, var, ne, _, _ = bn_fwd(a[pre + "post_expanded"], P[li + 7], P[li + 8], eps, False)
            a[pre + "norm_post_expanded.means"], a[pre + "norm_post_expanded.vars"], a[pre + "post_expanded_norm_vals"] = mu, var, ne
            if b["proj"]:
                a[pre + "transformed_residual"] = conv_fwd(x, P[li + 9], b["stride"])
                mu, var, sc, _, _ = bn_fwd(a[pre + "transformed_residual"], P[li + 10], P[li + 11], eps, False)
                a[pre + "norm_post_projection.means"], a[pre + "norm_post_projection.vars"] = mu, var
                a[pre + "post_projection_norm_vals"] = sc
                li += 12
            else:
                sc = x
                li += 9
            a[pre + "output_activated"], a[pre + "output"] = add_relu(ne, sc, keep_sum=True)
            x = a[pre + "output_activated"]
        a["final_conv_output_pooled"] = avgpool_fwd(x)
        a["linear_output"] = matmul(a["final_conv_output_pooled"], P[li])
        a["pred"] = softmax(a["linear_output"])
        return a["pred"]

    # -- backward, reference: resnet.cu:1777-2248 (+ resnet_clean.cu:2778 for the spatial BN)
    def backward(self):
        a, d, P, G, eps = self.act, {}, self.params, self.grads, self.eps
        self.dact = d
        nl = len(P)
        d["output_layer_deriv"] = ce_deriv(a["pred"], a["labels"])
        G[nl - 1] = matmul(a["final_conv_output_pooled"], d["output_layer_deriv"], ta=True)       # resnet.cu:1823
        d["final_conv_output_pooled"] = matmul(d["output_layer_deriv"], P[nl - 1], tb=True)        # resnet.cu:1830
        last = self.plan[-1]
        S_last = last["spatial"] // last["stride"]
        dOA = avgpool_bwd(d["final_conv_output_pooled"], S_last)
        # location index of each block's first tensor
        starts, li = [], 3
        for b in self.plan:
            starts.append(li)
            li += 12 if b["proj"] else 9
        for i in range(self.n_blocks - 1, -1, -1):
            b, li, pre = self.plan[i], starts[i], "b%d." % i
            d[pre + "output_activated"] = dOA
            x_in = a["init_convblock_input"] if i == 0 else a["b%d.output_activated" % (i - 1)]
            dO = relu_bwd(a[pre + "output"], dOA)                                                  # resnet.cu:1934
            d[pre + "output"] = dO
            if b["proj"]:
                dg, db, dXp = bn_bwd(a[pre + "transformed_residual"], P[li + 10], eps, a[pre + "norm_post_projection.means"],
                                     a[pre + "norm_post_projection.vars"], None, dO, False)
                G[li + 10], G[li + 11] = dg, db
                d[pre + "transformed_residual"] = dXp
                dBI = conv_dgrad(P[li + 9], dXp, b["spatial"], b["stride"])                        # resnet.cu:1991 (toAdd=false)
                G[li + 9] = conv_wgrad(x_in, dXp, b["proj_k"], b["stride"])
            else:
                dBI = dO.copy()                                                                    # resnet.cu:2003-2004
            dg, db, dXe = bn_bwd(a[pre + "post_expanded"], P[li + 7], eps, a[pre + "norm_post_expanded.means"],
                                 a[pre + "norm_post_expanded.vars"], None, dO, False)
            G[li + 7], G[li + 8] = dg, db
            d[pre + "post_expanded"] = dXe
            So = b["spatial"] // b["stride"]
            dYs = conv_dgrad(P[li + 6], dXe, So, 1)
            G[li + 6] = conv_wgrad(a[pre + "post_spatial_activated"], dXe, 1, 1)
            d[pre + "post_spatial_activated"] = dYs
            dg, db, dXs = bn_bwd(a[pre + "post_spatial"], P[li + 4], eps, a[pre + "norm_post_spatial.means"],
                                 a[pre + "norm_post_spatial.vars"], a[pre + "post_spatial_activated"], dYs, True)
            G[li + 4], G[li + 5] = dg, db
            d[pre + "post_spatial"] = dXs
            dYr = conv_dgrad(P[li + 3], dXs, b["spatial"], b["stride"])
            G[li + 3] = conv_wgrad(a[pre + "post_reduced_activated"], dXs, 3, b["stride"])
            d[pre + "post_reduced_activated"] = dYr
            dg, db, dXr = bn_bwd(a[pre + "post_reduced"], P[li + 1], eps, a[pre + "norm_post_reduced.means"],
                                 a[pre + "norm_post_reduced.vars"], a[pre + "post_reduced_activated"], dYr, True)
            G[li + 1], G[li + 2] = dg, db
            d[pre + "post_reduced"] = dXr
            dBI = conv_dgrad(P[li], dXr, b["spatial"], 1, din=dBI)                                 # resnet.cu:2157 (toAdd=true)
            G[li] = conv_wgrad(x_in, dXr, 1, 1)
            dOA = dBI
        d["init_convblock_input"] = dOA
        d["init_conv_activated"] = maxpool_bwd(a["max_inds"], dOA, a["init_conv_activated"].shape)
        dg, db, dX0 = bn_bwd(a["init_conv_applied"], P[1], eps, a["norm_init_conv.means"], a["norm_init_conv.vars"],
                             a["init_conv_activated"], d["init_conv_activated"], True)
        G[1], G[2] = dg, db
        d["init_conv_applied"] = dX0
        G[0] = conv_wgrad(a["images"], dX0, self.init_k, self.init_stride)                         # resnet.cu:2243 (no dgrad)
        return G

    # -- Adam, reference: resnet.cu:2910-2987
    def update(self):
        self.cur_b1 = float(np.float32(self.cur_b1) * np.float32(self.b1))
        self.cur_b2 = float(np.float32(self.cur_b2) * np.float32(self.b2))
        for i in range(len(self.params) - 1, -1, -1):
            g = np.ascontiguousarray(self.grads[i].reshape(self.shapes[i]), np.float32)
            adam(self.params[i], g, self.m[i], self.v[i], self.lr, self.wd, self.b1, self.b2, self.cur_b1, self.cur_b2,
                 self.eps)
        for i in range(len(self.grads)):
            self.grads[i] = np.zeros(self.shapes[i], np.float32)

    def loss_acc(self):
        return loss_acc(self.act["pred"], self.act["labels"])


def synthetic_batch(batch, input_dim, seed=1234, n_classes=1000):
    """SURVEY.md 8(d) synthetic inputs; ONE definition, shared with bench.py's product arm (which must not import the oracle):
    resnet_b200/synth.py (pure numpy)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("_rb_synth", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "resnet_b200", "synth.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.synthetic_batch(batch, input_dim, seed=seed, n_classes=n_classes)
