"""oracle/ref.py -- TEST INFRASTRUCTURE.  ctypes view of oracle/_ref/libref_*.so (the reference's own
translation units behind oracle/ref_harness.cu).  Needs a GPU; nothing here reads /root/reference."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
VARIANTS = {"naive": "libref_naive.so", "clean": "libref_clean.so", "cudnn": "libref_cudnn.so", "fast": "libref_fast.so",
            "fast_cached": "libref_fast_cached.so"}
f32p = C.POINTER(C.c_float)
i32p = C.POINTER(C.c_int)


def available(variant):
    return os.path.exists(os.path.join(_HERE, "_ref", VARIANTS[variant]))


def _f(a):
    return a.ctypes.data_as(f32p)


def _i(a):
    return a.ctypes.data_as(i32p)


class Ref:
    def __init__(self, variant):
        self.variant = variant
        self.lib = L = C.CDLL(os.path.join(_HERE, "_ref", VARIANTS[variant]))
        L.ref_create.restype = C.c_void_p
        L.ref_create.argtypes = [C.c_int, C.c_int, i32p, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float,
                                 C.c_float, C.c_ulonglong]
        L.ref_last_cuda_error.restype = C.c_char_p
        for n in ("ref_n_locations", "ref_location_size"):
            getattr(L, n).restype = C.c_int
        L.ref_n_locations.argtypes = [C.c_void_p]
        L.ref_location_size.argtypes = [C.c_void_p, C.c_int]
        L.ref_get_param.argtypes = [C.c_void_p, C.c_int, C.c_int, f32p]
        L.ref_set_param.argtypes = [C.c_void_p, C.c_int, C.c_int, f32p]
        L.ref_set_batch.argtypes = [C.c_void_p, f32p, i32p]
        for n in ("ref_forward", "ref_backward", "ref_update"):
            getattr(L, n).argtypes = [C.c_void_p]
        L.ref_get_pred.argtypes = [C.c_void_p, f32p]
        L.ref_time_steps.restype = C.c_double
        L.ref_time_steps.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
        L.ref_get_activation.restype = C.c_size_t
        L.ref_get_activation.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_void_p, C.c_size_t]
        self.h = None

    def create(self, input_dim, n_blocks, reductions, batch, output=1000, lr=1e-4, wd=0.0, b1=0.9, b2=0.999, eps=1e-7,
               seed=1234):
        red = np.asarray(reductions, np.int32)
        self.batch, self.output = batch, output
        self.h = self.lib.ref_create(input_dim, n_blocks, _i(red), batch, output, lr, wd, b1, b2, eps, seed)
        self.n_locations = self.lib.ref_n_locations(self.h)
        self.sizes = [self.lib.ref_location_size(self.h, i) for i in range(self.n_locations)]
        return self

    def cuda_error(self):
        return self.lib.ref_last_cuda_error().decode()

    def get_params(self, which=0):
        out = []
        for i, n in enumerate(self.sizes):
            a = np.empty(n, np.float32)
            self.lib.ref_get_param(self.h, which, i, _f(a))
            out.append(a)
        return out

    def set_params(self, arrays, which=0):
        for i, a in enumerate(arrays):
            a = np.ascontiguousarray(a, np.float32).reshape(-1)
            assert a.size == self.sizes[i]
            self.lib.ref_set_param(self.h, which, i, _f(a))

    def set_batch(self, images, labels):
        self.lib.ref_set_batch(self.h, _f(np.ascontiguousarray(images, np.float32)), _i(np.ascontiguousarray(labels, np.int32)))

    def forward(self):
        self.lib.ref_forward(self.h)
        p = np.empty((self.batch, self.output), np.float32)
        self.lib.ref_get_pred(self.h, _f(p))
        return p

    def backward(self):
        self.lib.ref_backward(self.h)

    def update(self):
        self.lib.ref_update(self.h)

    def time_steps(self, warmup, steps, e2e=False, forward_only=False):
        return self.lib.ref_time_steps(self.h, warmup, steps, int(e2e), int(forward_only))

    def activation(self, name, deriv=False, dtype=np.float32):
        n = self.lib.ref_get_activation(self.h, name.encode(), int(deriv), None, 0)
        if n == 0:
            return None
        a = np.empty(n, dtype)
        self.lib.ref_get_activation(self.h, name.encode(), int(deriv), a.ctypes.data_as(C.c_void_p), n)
        return a

    # ---- single-kernel entry points (naive variant only)
    def op_conv_fwd(self, x, w, stride):
        N, S, _, cin = x.shape
        cout, _, k, _ = w.shape
        out = np.empty((N, S // stride, S // stride, cout), np.float32)
        self.lib.ref_op_conv_fwd(_f(x), _f(w), S, k, cin, cout, stride, N, _f(out))
        return out

    def op_conv_bwd(self, x, w, dout, stride, din_base=None, want_din=True):
        N, S, _, cin = x.shape
        cout, _, k, _ = w.shape
        dw = np.empty_like(w)
        din = None
        if want_din:
            din = din_base.copy() if din_base is not None else np.zeros_like(x)
        self.lib.ref_op_conv_bwd(_f(x), _f(w), _f(dout), S, k, cin, cout, stride, N, int(din_base is not None),
                                 _f(din) if din is not None else None, _f(dw))
        return din, dw

    def op_bn_fwd(self, x, gamma, beta, eps, relu):
        N, S, _, Cc = x.shape
        means, vars_ = np.empty(Cc, np.float32), np.empty(Cc, np.float32)
        xhat, norm, act = np.empty_like(x), np.empty_like(x), np.empty_like(x)
        self.lib.ref_op_bn_fwd(_f(x), _f(gamma), _f(beta), S, Cc, N, C.c_float(eps), int(relu), _f(means), _f(vars_),
                               _f(xhat), _f(norm), _f(act))
        return means, vars_, xhat, norm, act

    def op_bn_bwd(self, x, gamma, beta, eps, relu, means, vars_, xhat, act, dy):
        N, S, _, Cc = x.shape
        dg, db, dx = np.empty(Cc, np.float32), np.empty(Cc, np.float32), np.empty_like(x)
        self.lib.ref_op_bn_bwd(_f(x), _f(gamma), _f(beta), S, Cc, N, C.c_float(eps), int(relu), _f(means), _f(vars_),
                               _f(xhat), _f(act), _f(dy), _f(dg), _f(db), _f(dx))
        return dg, db, dx

    def op_maxpool_fwd(self, x, k, stride):
        N, S, _, Cc = x.shape
        So = S // stride
        out, inds = np.empty((N, So, So, Cc), np.float32), np.empty((N, So, So, Cc), np.int32)
        self.lib.ref_op_maxpool_fwd(_f(x), k, stride, S, Cc, N, _i(inds), _f(out))
        return out, inds

    def op_adam(self, p, g, m, v, lr, wd, b1, b2, cur_b1, cur_b2, eps):
        self.lib.ref_op_adam(_f(p), _f(g), _f(m), _f(v), p.size, C.c_float(lr), C.c_float(wd), C.c_float(b1),
                             C.c_float(b2), C.c_float(cur_b1), C.c_float(cur_b2), C.c_float(eps))
