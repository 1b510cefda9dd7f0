"""One-off parity record at a batch size where BatchNorm is well populated (default 32): the full ResNet-50 step on the B200 path
(TF32 and bf16 storage) against the host oracle on the same cuRAND-initialised weights and synthetic batch.  Prints whole-tensor
rel-L2 of block outputs / logits, softmax max-abs, loss, argmax agreement, and gradient rel-L2 (profiles/r01_parity_batch32.txt).
The oracle needs ~0.2 GB of host memory and ~0.7 s per image."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
from resnet_b200 import api  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 32
red = [1 if i in (3, 7, 13) else 0 for i in range(16)]


def rel_l2(a, b):
    return float(np.linalg.norm(a.reshape(-1) - b.reshape(-1)) / max(1e-12, np.linalg.norm(b)))


img, lab = O.synthetic_batch(N, 224, seed=1234)
ref = api.Trainer(input_dim=224, n_blocks=16, reductions=red, batch=N, output=1000, lr=1e-4, seed=1234, dtype="tf32")
net = O.OracleNet(224, 16, red, batch=N, output=1000, lr=1e-4)
W = [w.reshape(s).copy() for w, s in zip(ref.get_params(0), net.shapes)]
ref.close()
names = ["init_convblock_input", "b0.output_activated", "b3.output_activated", "b7.output_activated", "b13.output_activated",
         "b15.output_activated", "linear_output"]
for damp in (None, 0.25):
    Wd = [w.copy() for w in W]
    if damp is not None:
        li = 3
        for b in net.plan:
            Wd[li + 7][:] = damp
            li += 12 if b["proj"] else 9
    net.set_params([w.copy() for w in Wd])
    t0 = time.time()
    opred = net.forward(img, lab)
    og = [g.copy() for g in net.backward()]
    oloss, _ = net.loss_acc()
    print("oracle: batch %d, expansion-BN gamma %s, forward + backward %.1f s on %d threads" % (N, "1 (reference init)" if damp is None else damp, time.time() - t0, O.num_threads()), flush=True)
    for dtype in ("tf32", "bf16"):
        t = api.Trainer(input_dim=224, n_blocks=16, reductions=red, batch=N, output=1000, lr=1e-4, seed=1234, dtype=dtype)
        t.set_params(Wd)
        t.set_batch(img, lab)
        pred = t.forward()
        errs = {nm: rel_l2(t.activation(nm), net.act[nm]) for nm in names}
        loss, _ = t.loss_accuracy()
        t.backward()
        tg = t.get_params(1)
        per = {i: rel_l2(g, r) for i, (g, r) in enumerate(zip(tg, og)) if len(net.shapes[i]) > 1}
        whole = rel_l2(np.concatenate([g.reshape(-1) for g in tg]), np.concatenate([r.reshape(-1) for r in og]))
        print("  %s: activations %s" % (dtype, " ".join("%s=%.1e" % (k.replace(".output_activated", ""), v) for k, v in errs.items())))
        print("  %s: softmax max-abs %.2e, argmax agreement %d/%d, loss %.6f vs %.6f, gradient vector rel-L2 %.2e, FC %.2e, worst weight tensor %.2e" %
              (dtype, float(np.abs(pred - opred).max()), int((pred.argmax(1) == opred.argmax(1)).sum()), N, loss, oloss, whole, per[max(per)], max(per.values())), flush=True)
        t.close()
