"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol the headers declare,
the ctypes struct mirrors match the C layout of include/resnet.h, and host-only logic (bucket planning) is right.
No compute calls: there is no GPU here."""
import ctypes as C
import os
import re
import subprocess

import pytest

from resnet_b200 import lib as rlib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def L():
    import __graft_entry__ as g
    if not os.path.exists(rlib.SO_PATH):
        g.build()
    return rlib.load()


def _declared(header):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return set(re.findall(r"\b([a-z_][a-z0-9_]*)\s*\(", src)) - {"defined", "sizeof"}


def test_library_exports_every_declared_symbol(L):
    for header, listed in (("resnet.h", rlib.RESNET_H_SYMBOLS), ("resnet_b200.h", rlib.RESNET_B200_H_SYMBOLS)):
        declared = {s for s in _declared(header) if s.startswith(("resnet_b200_", "init_", "forward_", "backwards_", "update_", "load_", "populate_", "dump_", "overwrite_"))}
        assert declared == set(listed), (header, declared ^ set(listed))
        for sym in declared:
            assert hasattr(L, sym), sym


def test_struct_layout_matches_c_header(tmp_path):
    """sizeof/offsetof from a C compile of include/resnet.h == the ctypes mirrors (drop-in ABI)."""
    names = ["Dims", "BatchNorm", "ConvBlock", "Params", "Cache_BatchNorm", "Activation_ConvBlock", "Activations", "ResNet",
             "Forward_Buffer", "Backprop_Buffer", "Batch", "Train_ResNet"]
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "resnet.h"', "int main(void){"]
    for n in names:
        cls = getattr(rlib, n)
        lines.append('printf("%s %%zu\\n", sizeof(%s));' % (n, n))
        for f, _ in cls._fields_:
            lines.append('printf("%s.%s %%zu\\n", offsetof(%s, %s));' % (n, f, n, f))
    lines.append("return 0;}")
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.check_call(["/usr/bin/gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    out = dict(l.split() for l in subprocess.check_output([str(exe)], text=True).splitlines())
    for n in names:
        cls = getattr(rlib, n)
        assert int(out[n]) == C.sizeof(cls), n
        for f, _ in cls._fields_:
            assert int(out["%s.%s" % (n, f)]) == getattr(cls, f).offset, (n, f)


def test_struct_fields_match_reference_header_names():
    """Field names and order follow the reference's resnet.h (recorded here from reference: resnet.h:176-215)."""
    assert [f for f, _ in rlib.Train_ResNet._fields_] == [
        "model", "cur_batch", "forward_buffer", "backprop_buffer", "learning_rate", "weight_decay", "base_mean_decay",
        "base_var_decay", "cur_mean_decay", "cur_var_decay", "eps", "batch_size", "n_epochs", "cur_dump_id", "cur_epoch",
        "loss_per_epoch", "accuracy_per_epoch", "init_loaded", "dump_dir"]
    assert [f for f, _ in rlib.Batch._fields_] == [
        "image_dim", "image_size", "n_images", "cur_shard_id", "cur_batch_in_shard", "shard_n_images", "full_shard_images",
        "full_shard_correct_classes", "images_float_cpu", "images", "correct_classes_cpu", "correct_classes"]


def test_init_dimensions_is_host_only(L):
    red = (C.c_int * 4)(0, 1, 0, 1)
    d = L.init_dimensions(32, 7, 64, 2, 3, 2, 4, red, 1024, 10).contents
    assert (d.input, d.init_kernel_dim, d.n_conv_blocks, d.final_depth, d.output) == (32, 7, 4, 1024, 10)
    assert [d.is_block_spatial_reduction[i] for i in range(4)] == [0, 1, 0, 1]


def test_dp_bucket_plan(L):
    """buckets cover the gradient arena exactly once, deepest first, cut at block boundaries; the stem rides last."""
    from oracle import oracle as O
    shapes = O.param_shapes(224, 16, [1 if i in (3, 7, 13) else 0 for i in range(16)])
    offs, off = [], 0
    for s in shapes:
        offs.append(off)
        n = 1
        for v in s:
            n *= v
        off += (n + 63) // 64 * 64
    total = off
    plan = O.block_plan(224, 16, [1 if i in (3, 7, 13) else 0 for i in range(16)])
    first, li = [], 3
    for b in plan:
        first.append(li)
        li += 12 if b["proj"] else 9
    n = 64
    o_off, o_len, o_fb = (C.c_longlong * n)(), (C.c_longlong * n)(), (C.c_int * n)()
    cnt = L.resnet_b200_dp_plan((C.c_longlong * len(offs))(*offs), len(offs), total, (C.c_int * 16)(*first), 16, (32 << 20) // 4,
                                o_off, o_len, o_fb, n)
    b = [(o_off[i], o_len[i], o_fb[i]) for i in range(cnt)]
    assert cnt >= 3
    assert b[0][0] + b[0][1] == total and b[-1][0] == 0 and b[-1][2] == -1
    for (o1, l1, f1), (o2, l2, f2) in zip(b, b[1:]):
        assert o2 + l2 == o1                      # contiguous, descending
        assert f2 < f1 or f2 == -1                # issue order follows backward's block order
    assert sum(x[1] for x in b) == total
    assert all(x[1] * 4 >= (32 << 20) for x in b[:-1])
