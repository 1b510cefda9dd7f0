"""How long does the GPU idle between forward_pass and backwards_pass when the host waits for pred_cpu (the reference's contract)?
Same 20 steps with resnet_b200_set_pred_copy 1 (default: D2H + stream sync inside forward_pass) and 0 (no host sync inside the step)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from resnet_b200 import api, synth as O
L = api.L()
for dtype in ("tf32", "bf16"):
    red = [1 if i in (3, 7, 13) else 0 for i in range(16)]
    t = api.Trainer(input_dim=224, n_blocks=16, reductions=red, batch=256, output=1000, lr=1e-4, seed=1234, device=0, dtype=dtype)
    img, lab = O.synthetic_batch(256, 224, seed=1234)
    t.set_batch(img, lab)
    for mode in (1, 0, 1, 0):
        L.resnet_b200_set_pred_copy(t.t, mode)
        for _ in range(5):
            L.forward_pass(t.t); L.backwards_pass(t.t); L.update_parameters(t.t)
        t.sync()
        L.resnet_b200_timer_begin(t.t)
        for _ in range(20):
            L.forward_pass(t.t); L.backwards_pass(t.t); L.update_parameters(t.t)
        ms = L.resnet_b200_timer_end_ms(t.t)
        print("%s pred_copy=%d: %.3f ms per step" % (dtype, mode, ms / 20), flush=True)
    t.close()
