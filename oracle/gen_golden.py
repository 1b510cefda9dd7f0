"""oracle/gen_golden.py -- TEST INFRASTRUCTURE.  Run on a B200 box:

    python oracle/gen_golden.py gpurun_out/reference_b200.npz

Executes the reference's own compiled kernels / entry points (oracle/_ref/libref_*.so, built from
/root/reference by oracle/Makefile) on the deterministic cases of oracle/golden_cases.py and stores their
outputs.  The result is committed as tests/golden/reference_b200.npz; tests replay the same cases through
the oracle (CPU, `-m "not gpu"`) and through libresnet_b200.so (`-m gpu`).

Every stage runs in its own subprocess: the reference exit(1)s / dereferences NULL FILE*s on its error paths
(reference: resnet.cu:2896-2899, 2285-2314), and one variant failing must not lose the others.
"""
import os
import subprocess
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import golden_cases as G  # noqa: E402
from oracle import oracle as O  # noqa: E402
from oracle.ref import Ref, available  # noqa: E402


def stage_ops(out):
    naive = Ref("naive")
    for i in range(len(G.CONV_CASES)):
        S, k, cin, cout, stride, N = G.CONV_CASES[i]
        x, w, dy, base = G.conv_inputs(i)
        out["conv%d.y" % i] = naive.op_conv_fwd(x, w, stride)
        din, dw = naive.op_conv_bwd(x, w, dy, stride)
        out["conv%d.din" % i], out["conv%d.dw" % i] = din, dw
        din_add, _ = naive.op_conv_bwd(x, w, dy, stride, din_base=base)
        out["conv%d.din_add" % i] = din_add
    for i in range(len(G.BN_CASES)):
        x, g, b, dy, relu = G.bn_inputs(i)
        mu, var, xh, nv, act = naive.op_bn_fwd(x, g, b, 1e-7, relu)
        out["bn%d.means" % i], out["bn%d.vars" % i], out["bn%d.xhat" % i] = mu, var, xh
        out["bn%d.normalized" % i], out["bn%d.activated" % i] = nv, act
        dg, db, dx = naive.op_bn_bwd(x, g, b, 1e-7, relu, mu, var, xh, act, dy)
        out["bn%d.dgamma" % i], out["bn%d.dbeta" % i], out["bn%d.dx" % i] = dg, db, dx
    mp, inds = naive.op_maxpool_fwd(G.maxpool_input(), 3, 2)
    out["maxpool.out"], out["maxpool.inds"] = mp, inds
    p, g1, g2 = G.adam_inputs()
    m, v = np.zeros_like(p), np.zeros_like(p)
    naive.op_adam(p, g1, m, v, 1e-3, 0.0, 0.9, 0.999, 0.9, 0.999, 1e-7)
    out["adam.p1"], out["adam.m1"], out["adam.v1"] = p.copy(), m.copy(), v.copy()
    naive.op_adam(p, g2, m, v, 1e-3, 0.01, 0.9, 0.999, 0.81, 0.998001, 1e-7)
    out["adam.p2"], out["adam.m2"], out["adam.v2"] = p.copy(), m.copy(), v.copy()
    return "ops: " + naive.cuda_error()


def stage_net(out, tag, variant):
    cfg = {"mini": G.MINI, "mini4": G.MINI4, "mini5": G.MINI5}[tag]
    shapes = O.param_shapes(cfg["input_dim"], cfg["n_blocks"], cfg["reductions"], output=cfg["output"])
    W = G.mini_weights(shapes)
    img, lab = G.mini_batch(cfg)
    r = Ref(variant).create(seed=1234, **cfg)
    r.n_locations = len(shapes)  # the reference over-counts locations unless there are exactly 4 projections (resnet.cu:819)
    r.sizes = r.sizes[:len(shapes)]
    if tag == "mini" and variant == "naive":
        out["mini.curand_init"] = np.stack([G.summary(a) for a in r.get_params()])  # the reference's own cuRAND init
    r.set_params(W)
    r.set_batch(img, lab)
    key = "%s.%s" % (tag, variant)
    out[key + ".pred"] = r.forward()
    if variant == "naive":
        for nm in G.ACT_NAMES_FWD:
            out[key + ".act." + nm] = G.summary(r.activation(nm))
        out[key + ".max_inds"] = r.activation("max_inds", dtype=np.int32)
        for bi in range(cfg["n_blocks"]):
            for f in G.BLOCK_FIELDS_FWD:
                a = r.activation("b%d.%s" % (bi, f))
                if a is not None:
                    out[key + ".act.b%d.%s" % (bi, f)] = G.summary(a)
        return key + " fwd: " + r.cuda_error()  # resnet.cu's block backward is incomplete (resnet.cu:2060-2083)
    r.backward()
    grads = r.get_params(1)
    if not all(np.isfinite(g).all() for g in grads):
        return key + ": non-finite gradients from the reference, step not recorded"
    out[key + ".grads"] = np.stack([G.summary(a) for a in grads])
    if variant == "cudnn":
        for nm in ("init_convblock_input", "init_conv_applied", "b0.post_reduced", "b1.post_expanded", "b1.transformed_residual",
                   "b1.post_spatial", "b0.output_activated"):
            a = r.activation(nm, deriv=True)
            if a is not None:
                out[key + ".dact." + nm] = G.summary(a)
    r.update()
    out[key + ".params1"] = np.stack([G.summary(a) for a in r.get_params(0)])
    out[key + ".m1"] = np.stack([G.summary(a) for a in r.get_params(2)])
    r.set_batch(img, lab)  # the reference zeroes grads and the batch after a step; feed the batch again
    out[key + ".pred2"] = r.forward()
    r.backward()
    r.update()
    out[key + ".params2"] = np.stack([G.summary(a) for a in r.get_params(0)])
    return key + " step: " + r.cuda_error()


def run_stage(stage, part_path):
    out = {}
    if stage == "ops":
        note = stage_ops(out)
    else:
        _, tag, variant = stage.split(":")
        note = stage_net(out, tag, variant)
    out["note"] = np.array(note)
    np.savez_compressed(part_path, **out)


def main(out_path):
    os.makedirs(os.path.dirname(os.path.abspath(out_path)), exist_ok=True)
    stages = ["ops", "net:mini:naive", "net:mini4:naive"] + ["net:mini5:%s" % v for v in ("naive", "clean", "cudnn") if available(v)]
    merged, notes = {}, []
    for st in stages:
        part = out_path + "." + st.replace(":", "_") + ".part.npz"
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--stage", st, part], capture_output=True, text=True, timeout=240)
            rc = r.returncode
        except subprocess.TimeoutExpired:
            rc = "timeout"
        if os.path.exists(part):
            d = np.load(part)
            for k in d.files:
                if k == "note":
                    notes.append(str(d[k]))
                else:
                    merged[k] = d[k]
            os.remove(part)
        if rc != 0:
            notes.append("%s: subprocess rc=%s" % (st, rc))
    merged["notes"] = np.array("; ".join(notes))
    np.savez_compressed(out_path, **merged)
    print("wrote", out_path, "keys:", len(merged))
    print(merged["notes"])


if __name__ == "__main__":
    if len(sys.argv) > 3 and sys.argv[1] == "--stage":
        run_stage(sys.argv[2], sys.argv[3])
    else:
        main(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/reference_b200.npz")
