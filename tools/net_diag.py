"""Per-tensor error report of the whole step vs the oracle, for both conv modes (bring-up aid; prints, never asserts).
usage: python tools/net_diag.py [MINI|MINI4] > report.txt"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import golden_cases as G  # noqa: E402
from oracle import oracle as O  # noqa: E402
from resnet_b200 import api  # noqa: E402


def rel_max(a, b):
    return float(np.abs(a.reshape(-1) - b.reshape(-1)).max() / max(1e-9, np.abs(b).max()))


def rel_l2(a, b):
    return float(np.linalg.norm(a.reshape(-1) - b.reshape(-1)) / max(1e-9, np.linalg.norm(b)))


def main():
    cfg = getattr(G, sys.argv[1] if len(sys.argv) > 1 else "MINI")
    shapes = O.param_shapes(cfg["input_dim"], cfg["n_blocks"], cfg["reductions"], output=cfg["output"])
    W = G.mini_weights(shapes)
    img, lab = G.mini_batch(cfg)
    net = O.OracleNet(cfg["input_dim"], cfg["n_blocks"], cfg["reductions"], cfg["batch"], output=cfg["output"], lr=cfg["lr"])
    net.set_params([w.copy() for w in W])
    opred = net.forward(img, lab)
    og = [g.copy() for g in net.backward()]
    for mode in ("simt", "tc"):
        os.environ["RESNET_B200_CONV"] = mode
        os.environ["RESNET_B200_KEEP_ALL"] = "1"
        t = api.Trainer(input_dim=cfg["input_dim"], n_blocks=cfg["n_blocks"], reductions=cfg["reductions"], batch=cfg["batch"],
                        output=cfg["output"], lr=cfg["lr"])
        t.set_params(W)
        t.set_batch(img, lab)
        pred = t.forward()
        print("==== mode %s  tensor cores: %s  pred rel_max %.3e argmax_eq %s" % (mode, t.uses_tensor_cores(), rel_max(pred, opred),
                                                                                  (pred.argmax(1) == opred.argmax(1)).all()))
        for nm in sorted(net.act):
            if nm in ("images", "labels", "pred", "max_inds"):
                continue
            got = t.activation(nm)
            if got is not None:
                print("  act  %-36s rel_max %.3e" % (nm, rel_max(got, net.act[nm])))
        t.backward()
        for i, (g, r) in enumerate(zip(t.get_params(1), og)):
            print("  grad %3d %-20s rel_l2 %.3e  |ref| %.3e" % (i, shapes[i], rel_l2(g, r), float(np.linalg.norm(r))))
        for nm in sorted(net.dact):
            try:
                got = t.activation(nm, deriv=True)
            except Exception:
                got = None
            if got is not None:
                print("  dact %-36s rel_l2 %.3e" % (nm, rel_l2(got, net.dact[nm])))
        t.close()


if __name__ == "__main__":
    main()
