"""Host-side mirror of the reference's driver API over libresnet_b200.so.

The calls are the reference's own (`init_dimensions`, `init_resnet`, `init_general_batch`, `init_trainer`,
`forward_pass`, `backwards_pass`, `update_parameters`; reference: resnet.cu:3259-3402) -- this module only adds
numpy <-> device conveniences so that tests and bench.py read like the reference's main loop.
"""
import ctypes as C

import numpy as np

from . import lib as _lib

f32p = _lib.f32p
i32p = _lib.i32p


def L():
    return _lib.load()


def check():
    err = L().resnet_b200_last_error().decode()
    if err:
        raise RuntimeError("libresnet_b200: " + err)


def to_bf16(a):
    """float32 -> bf16 bits (uint16), round to nearest even: what cvt.rn.bf16.f32 does on the device"""
    u = np.ascontiguousarray(a, np.float32).view(np.uint32).astype(np.uint64)
    return ((u + 0x7FFF + ((u >> 16) & 1)) >> 16).astype(np.uint16)


def from_bf16(u):
    return (np.ascontiguousarray(u, np.uint16).astype(np.uint32) << 16).view(np.float32)


def bf16_round(a):
    """float32 values rounded to the nearest bf16 (still float32)"""
    return from_bf16(to_bf16(a)).reshape(np.shape(a))


class _OpDtype:
    """single-operator calls inside this context read / write bf16 activation tensors"""

    def __init__(self, bf16):
        self.bf16 = bool(bf16)

    def __enter__(self):
        L().resnet_b200_set_op_dtype(int(self.bf16))
        return self

    def __exit__(self, *a):
        L().resnet_b200_set_op_dtype(0)

    def up(self, arr):
        """host fp32 activation tensor -> device buffer in the op dtype"""
        return DevBuf(to_bf16(arr) if self.bf16 else np.ascontiguousarray(arr, np.float32))

    def out(self, n):
        return DevBuf(nbytes=(2 if self.bf16 else 4) * int(n), zero=True)

    def down(self, buf, shape):
        return from_bf16(buf.get(shape, np.uint16)).reshape(shape) if self.bf16 else buf.get(shape)


class DevBuf:
    """A device allocation owned by Python (single-operator tests)."""

    def __init__(self, arr=None, nbytes=None, zero=False):
        self.nbytes = int(arr.nbytes if arr is not None else nbytes)
        self.ptr = L().resnet_b200_malloc(self.nbytes)
        if arr is not None:
            a = np.ascontiguousarray(arr)
            L().resnet_b200_memcpy_h2d(self.ptr, a.ctypes.data_as(C.c_void_p), self.nbytes)
        elif zero:
            L().resnet_b200_memset(self.ptr, 0, self.nbytes)

    def get(self, shape, dtype=np.float32):
        out = np.empty(shape, dtype)
        assert out.nbytes <= self.nbytes
        L().resnet_b200_memcpy_d2h(out.ctypes.data_as(C.c_void_p), self.ptr, out.nbytes)
        return out

    def free(self):
        if self.ptr:
            L().resnet_b200_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def d2h(ptr, n, dtype=np.float32):
    """copy n elements from a device pointer (ctypes pointer or int)"""
    out = np.empty(int(n), dtype)
    addr = C.cast(ptr, C.c_void_p)
    L().resnet_b200_memcpy_d2h(out.ctypes.data_as(C.c_void_p), addr, out.nbytes)
    return out


def h2d(ptr, arr):
    a = np.ascontiguousarray(arr)
    L().resnet_b200_memcpy_h2d(C.cast(ptr, C.c_void_p), a.ctypes.data_as(C.c_void_p), a.nbytes)


class Trainer:
    """ResNet trainer driven through the reference's entry points."""

    def __init__(self, input_dim=224, n_blocks=16, reductions=None, batch=32, output=1000, lr=1e-4, wd=0.0, b1=0.9, b2=0.999,
                 eps=1e-7, seed=1234, init_filters=64, device=None, shard_n_images=None, dtype=None, selfcheck=False):
        """dtype: None = follow $RESNET_B200_DTYPE, "tf32" = fp32 tensors / TF32 MMAs, "bf16" = bf16 tensors (include/resnet_b200.h)
        selfcheck: every tensor-core convolution launch is re-derived in place by the fp32 SIMT checker (resnet_b200_selfcheck)"""
        lib = L()
        if device is not None:
            lib.resnet_b200_set_device(int(device))
        if reductions is None:
            reductions = [1 if i in (3, 7, 13) else 0 for i in range(n_blocks)]
        self.input_dim, self.n_blocks, self.batch, self.output = input_dim, n_blocks, batch, output
        self.reductions = list(reductions)
        self._red = (C.c_int * n_blocks)(*self.reductions)
        final_depth = 4 * init_filters * (2 ** sum(self.reductions))
        self.final_depth = final_depth
        # reference: resnet.cu:3245-3296
        self.dims = lib.init_dimensions(input_dim, 7, init_filters, 2, 3, 2, n_blocks, self._red, final_depth, output)
        self.gen = lib.resnet_b200_rng_create(seed)
        self.model = lib.init_resnet(self.dims, self.gen)
        self.batch_struct = lib.init_general_batch(batch, input_dim * input_dim * 3, input_dim, shard_n_images or batch)
        self._dump_dir = b"resnet_b200"
        lib.resnet_b200_set_dtype({None: -1, "tf32": 0, "fp32": 0, "bf16": 1}[dtype])
        lib.resnet_b200_selfcheck(int(bool(selfcheck)))
        self.t = lib.init_trainer(self.model, self.batch_struct, batch, lr, wd, b1, b2, eps, 1, self._dump_dir)
        lib.resnet_b200_selfcheck(0)
        lib.resnet_b200_set_dtype(-1)
        check()
        self.bf16 = lib.resnet_b200_trainer_dtype(self.t) == 1
        P = self.model.contents.params.contents
        self.n_locations = P.n_locations
        self.sizes = [P.sizes[i] for i in range(P.n_locations)]

    # ---- parameters / optimizer state, in locations[] order (reference: resnet.h:85-87)
    def _tree(self, which):
        bb = self.t.contents.backprop_buffer.contents
        return [self.model.contents.params, bb.param_derivs, bb.prev_means, bb.prev_vars][which].contents

    def get_params(self, which=0):
        self.sync()
        tr = self._tree(which)
        return [d2h(tr.locations[i], self.sizes[i]) for i in range(self.n_locations)]

    def set_params(self, arrays, which=0):
        self.sync()
        tr = self._tree(which)
        for i, a in enumerate(arrays):
            a = np.ascontiguousarray(a, np.float32).reshape(-1)
            assert a.size == self.sizes[i], (i, a.size, self.sizes[i])
            h2d(tr.locations[i], a)

    # ---- batch
    def set_batch(self, images, labels):
        """blocking copy into cur_batch->images / correct_classes (what load_new_batch does, reference: resnet.cu:1315-1316)"""
        self.sync()
        b = self.batch_struct.contents
        h2d(b.images, np.ascontiguousarray(images, np.float32))
        h2d(b.correct_classes, np.ascontiguousarray(labels, np.int32))
        C.memmove(b.correct_classes_cpu, np.ascontiguousarray(labels, np.int32).ctypes.data, 4 * self.batch)

    def load_new_batch(self):
        """reference: resnet.cu:1235 -- next batch of the current shard ($RESNET_B200_SHARD_DIR/%03d.images|labels)"""
        L().load_new_batch(self.t, None, self.batch_struct)
        check()

    # ---- the reference's step
    def forward(self):
        L().forward_pass(self.t)
        check()
        return np.ctypeslib.as_array(self.t.contents.forward_buffer.contents.pred_cpu, shape=(self.batch, self.output)).copy()

    def backward(self):
        L().backwards_pass(self.t)
        check()

    def update(self):
        L().update_parameters(self.t)
        check()

    def sync(self):
        L().resnet_b200_trainer_sync(self.t)
        check()

    # ---- dumps / checkpoints in the reference's directory format (reference: resnet.cu:2755, 2778, 2821); root = $RESNET_B200_DUMP_ROOT
    def dump(self, dump_id, special_dir="resnet_b200"):
        L().dump_trainer(int(dump_id), self.t, special_dir.encode())
        check()

    def restore(self, dump_id, special_dir="resnet_b200"):
        L().overwrite_trainer_hyperparams(self.t, int(dump_id), special_dir.encode())
        L().overwrite_model_params(self.t, int(dump_id), special_dir.encode())
        check()

    def loss_accuracy(self):
        ls, nw = C.c_float(), C.c_int()
        L().resnet_b200_loss_accuracy(self.t, C.byref(ls), C.byref(nw))
        return ls.value, nw.value

    def set_pred_copy(self, on):
        """False: forward_pass neither copies pred to pred_cpu nor synchronises (epoch_stats / fetch_pred read results on demand)"""
        L().resnet_b200_set_pred_copy(self.t, int(bool(on)))
        check()

    def forward_async(self):
        L().forward_pass(self.t)

    def fetch_pred(self):
        L().resnet_b200_fetch_pred(self.t)
        check()
        return np.ctypeslib.as_array(self.t.contents.forward_buffer.contents.pred_cpu, shape=(self.batch, self.output)).copy()

    def epoch_stats(self, reset=False):
        """(loss sum, wrong predictions, images) accumulated on the device since the last reset (reference: resnet.cu:3363-3412)"""
        ls, nw, ni = C.c_double(), C.c_longlong(), C.c_longlong()
        L().resnet_b200_epoch_stats(self.t, C.byref(ls), C.byref(nw), C.byref(ni), int(reset))
        check()
        return ls.value, nw.value, ni.value

    def selfcheck_report(self):
        """{family: (worst max|diff| / max|ref|, launches checked, where)} for fprop / dgrad / wgrad (Trainer(selfcheck=True))"""
        out = {}
        for fam, name in enumerate(("fprop", "dgrad", "wgrad")):
            w, n, where = C.c_float(), C.c_longlong(), C.create_string_buffer(128)
            L().resnet_b200_selfcheck_read(self.t, fam, C.byref(w), C.byref(n), where, 128)
            check()
            out[name] = (w.value, n.value, where.value.decode())
        return out

    def uses_tensor_cores(self):
        return bool(L().resnet_b200_uses_tensor_cores(self.t))

    # ---- named activations (same names as oracle.OracleNet.act / reference struct fields)
    def activation(self, name, deriv=False, dtype=np.float32):
        self.sync()
        t = self.t.contents
        A = (t.backprop_buffer.contents.activation_derivs if deriv else t.forward_buffer.contents.activations).contents
        N, d = self.batch, self.dims.contents
        S1 = d.input // d.init_conv_stride
        S2 = S1 // d.init_maxpool_stride
        F = d.init_conv_filters
        top = {"init_conv_applied": (A.init_conv_applied, N * S1 * S1 * F), "init_conv_activated": (A.init_conv_activated, N * S1 * S1 * F),
               "init_convblock_input": (A.init_convblock_input, N * S2 * S2 * F), "max_inds": (A.max_inds, N * S2 * S2 * F),
               "final_conv_output_pooled": (A.final_conv_output_pooled, N * d.final_depth), "linear_output": (A.linear_output, N * d.output)}
        if name in top:
            ptr, n = top[name]
        elif name.startswith("norm_init_conv."):
            ptr, n = getattr(A.norm_init_conv.contents, name.split(".")[1]), F
        else:
            bi, field = name.split(".", 1)
            b = A.activation_conv_blocks[int(bi[1:])].contents
            s_in, s_out = b.incoming_spatial_dim, b.incoming_spatial_dim // b.stride
            sizes = {"post_reduced": N * s_in * s_in * b.reduced_depth, "post_reduced_activated": N * s_in * s_in * b.reduced_depth,
                     "post_spatial": N * s_out * s_out * b.reduced_depth, "post_spatial_activated": N * s_out * s_out * b.reduced_depth}
            if "." in field:
                cache, f2 = field.split(".")
                cb = getattr(b, cache)
                if not cb:
                    return None
                ptr, n = getattr(cb.contents, f2), cb.contents.feature_size if f2 in ("means", "vars") else cb.contents.input_size
            else:
                ptr, n = getattr(b, field), sizes.get(field, N * s_out * s_out * b.expanded_depth)
        if not ptr:
            return None
        fp32_always = name in ("max_inds", "final_conv_output_pooled", "linear_output") or name.endswith((".means", ".vars"))
        if self.bf16 and not fp32_always:
            return from_bf16(d2h(ptr, n, np.uint16))
        return d2h(ptr, n, dtype)

    def close(self):
        if self.t:
            L().resnet_b200_destroy_trainer(self.t)
            self.t = None


# ------------------------------------------------------------------------------------------- single operators
def conv_forward(x, w, stride, impl=0, dtype="f32"):
    N, S, _, cin = x.shape
    cout, _, k, _ = w.shape
    with _OpDtype(dtype == "bf16") as T:
        dx = DevBuf(np.ascontiguousarray(x, np.float32)) if cin == 3 else T.up(x)  # the stem reads the fp32 batch in both modes
        dw = DevBuf(w)
        dy = T.out(N * (S // stride) ** 2 * cout)
        L().resnet_b200_conv_forward(S, k, cin, cout, stride, N, dx.ptr, dw.ptr, dy.ptr, impl)
        check()
        return T.down(dy, (N, S // stride, S // stride, cout))


def conv_backward(x, w, dy, stride, din_base=None, want_din=True, impl=0, dtype="f32"):
    N, S, _, cin = x.shape
    cout, _, k, _ = w.shape
    with _OpDtype(dtype == "bf16") as T:
        bx = DevBuf(np.ascontiguousarray(x, np.float32)) if cin == 3 else T.up(x)
        bw, bdy = DevBuf(w), T.up(dy)
        bdw = DevBuf(nbytes=w.nbytes, zero=True)
        bdin = None
        if want_din:
            bdin = T.up(din_base) if din_base is not None else T.out(x.size)
        L().resnet_b200_conv_backward(S, k, cin, cout, stride, N, int(din_base is not None), bx.ptr, bw.ptr, bdy.ptr,
                                      bdin.ptr if bdin else None, bdw.ptr, impl)
        check()
        return (T.down(bdin, x.shape) if bdin else None), bdw.get(w.shape)


def batchnorm_forward(x, gamma, beta, eps, relu, residual=None, round_tf32=False, dtype="f32"):
    N, S, _, Cc = x.shape
    with _OpDtype(dtype == "bf16") as T:
        bx, bg, bb = T.up(x), DevBuf(gamma), DevBuf(beta)
        bm, bv, by = DevBuf(nbytes=4 * Cc), DevBuf(nbytes=4 * Cc), T.out(x.size)
        br = T.up(residual) if residual is not None else None
        L().resnet_b200_batchnorm_forward(S, Cc, N, eps, bx.ptr, bg.ptr, bb.ptr, bm.ptr, bv.ptr, by.ptr, int(relu), br.ptr if br else None,
                                          int(round_tf32))
        check()
        return bm.get((Cc,)), bv.get((Cc,)), T.down(by, x.shape)


def batchnorm_backward(x, gamma, eps, means, vars_, activated, dy, relu, dtype="f32"):
    N, S, _, Cc = x.shape
    with _OpDtype(dtype == "bf16") as T:
        bx, bg, bm, bv, ba, bdy = T.up(x), DevBuf(gamma), DevBuf(means), DevBuf(vars_), T.up(activated), T.up(dy)
        bdg, bdb, bdx = DevBuf(nbytes=4 * Cc), DevBuf(nbytes=4 * Cc), T.out(x.size)
        L().resnet_b200_batchnorm_backward(S, Cc, N, eps, bx.ptr, bg.ptr, bm.ptr, bv.ptr, ba.ptr, bdy.ptr, bdg.ptr, bdb.ptr, bdx.ptr, int(relu))
        check()
        return bdg.get((Cc,)), bdb.get((Cc,)), T.down(bdx, x.shape)


def maxpool_forward(x, k, stride, dtype="f32"):
    N, S, _, Cc = x.shape
    So = S // stride
    with _OpDtype(dtype == "bf16") as T:
        bx, bi, bo = T.up(x), DevBuf(nbytes=4 * N * So * So * Cc), T.out(N * So * So * Cc)
        L().resnet_b200_maxpool_forward(bx.ptr, k, stride, S, Cc, N, bi.ptr, bo.ptr)
        check()
        return T.down(bo, (N, So, So, Cc)), bi.get((N, So, So, Cc), np.int32)


def maxpool_backward(inds, dout, in_shape, k, stride, dtype="f32"):
    N, S, _, Cc = in_shape
    with _OpDtype(dtype == "bf16") as T:
        bi, bd = DevBuf(inds), T.up(dout)
        bo = T.out(int(np.prod(in_shape)))
        L().resnet_b200_maxpool_backward(bi.ptr, bd.ptr, k, S, stride, Cc, N, bo.ptr)
        check()
        return T.down(bo, tuple(in_shape))


def avgpool_forward(x, dtype="f32"):
    N, S, _, Cc = x.shape
    with _OpDtype(dtype == "bf16") as T:
        bx, bo = T.up(x), DevBuf(nbytes=4 * N * Cc)
        L().resnet_b200_avgpool_forward(bx.ptr, S, Cc, N, bo.ptr)
        check()
        return bo.get((N, Cc))


def avgpool_backward(dp, S, dtype="f32"):
    N, Cc = dp.shape
    with _OpDtype(dtype == "bf16") as T:
        bd, bo = DevBuf(dp), T.out(N * S * S * Cc)
        L().resnet_b200_avgpool_backward(bd.ptr, Cc, N, S, bo.ptr)
        check()
        return T.down(bo, (N, S, S, Cc))


def matmul(A, B, ta=False, tb=False):
    m, k = (A.shape[1], A.shape[0]) if ta else A.shape
    n = B.shape[0] if tb else B.shape[1]
    ba, bb, bo = DevBuf(A), DevBuf(B), DevBuf(nbytes=4 * m * n)
    L().resnet_b200_matmul(ba.ptr, bb.ptr, m, k, n, int(ta), int(tb), bo.ptr)
    check()
    return bo.get((m, n))


def softmax_ce(logits, labels):
    N, Lc = logits.shape
    bl, bi = DevBuf(logits), DevBuf(np.ascontiguousarray(labels, np.int32))
    bp, bd = DevBuf(nbytes=logits.nbytes), DevBuf(nbytes=logits.nbytes)
    L().resnet_b200_softmax_ce(bl.ptr, bi.ptr, N, Lc, bp.ptr, bd.ptr)
    check()
    return bp.get(logits.shape), bd.get(logits.shape)


def adam(p, g, m, v, lr, wd, b1, b2, cur_b1, cur_b2, eps):
    n = p.size
    pad = (-n) % 4
    bufs = [DevBuf(np.concatenate([a.reshape(-1), np.zeros(pad, np.float32)])) for a in (p, g, m, v)]
    L().resnet_b200_adam(bufs[0].ptr, bufs[1].ptr, bufs[2].ptr, bufs[3].ptr, n + pad, lr, wd, b1, b2, cur_b1, cur_b2, eps)
    check()
    return [b.get((n + pad,))[:n].reshape(p.shape) for b in bufs]
