"""GPU test of the dump / checkpoint format through the reference's entry points dump_trainer,
overwrite_trainer_hyperparams, overwrite_model_params (reference: resnet.cu:2755, 2778, 2821): file layout, fp32 contents equal to
the device buffers (bit-exact), and a restore that continues training bit-identically."""
import os

import numpy as np
import pytest

from oracle import golden_cases as G
from oracle import oracle as O

pytestmark = pytest.mark.gpu


def make(cfg, dtype, keep_all):
    from resnet_b200 import api
    os.environ["RESNET_B200_KEEP_ALL"] = "1" if keep_all else "0"
    try:
        return api.Trainer(input_dim=cfg["input_dim"], n_blocks=cfg["n_blocks"], reductions=cfg["reductions"], batch=cfg["batch"],
                           output=cfg["output"], lr=cfg["lr"], dtype=dtype)
    finally:
        os.environ.pop("RESNET_B200_KEEP_ALL", None)


@pytest.mark.parametrize("dtype,keep_all", [("tf32", True), ("bf16", False)])
def test_dump_layout_and_restore(tmp_path, dtype, keep_all):
    os.environ["RESNET_B200_DUMP_ROOT"] = str(tmp_path)
    try:
        cfg = G.MINI
        shapes = O.param_shapes(cfg["input_dim"], cfg["n_blocks"], cfg["reductions"], output=cfg["output"])
        t = make(cfg, dtype, keep_all)
        t.set_params(G.mini_weights(shapes))
        img, lab = G.mini_batch(cfg)
        for _ in range(2):
            t.set_batch(img, lab)
            t.forward()
            t.backward()
            t.update()
        t.set_batch(img, lab)
        pred = t.forward()
        t.backward()          # gradients present in the dump; the update comes after it
        t.dump(7, "run_a")
        root = tmp_path / "run_a" / "00000007"
        # ---- parameters, gradients, Adam state: one %03d.buffer per location, fp32, equal to the device arenas
        for which, name in enumerate(["model_params", "gradients", "means", "vars"]):
            dev = t.get_params(which)
            for i, a in enumerate(dev):
                f = np.fromfile(root / name / ("%03d.buffer" % i), np.float32)
                np.testing.assert_array_equal(f, a)
        # ---- activations in the reference's file names (reference: resnet.cu:2351-2680)
        act = root / "activations"
        np.testing.assert_array_equal(np.fromfile(act / "input.buffer", np.float32), img.reshape(-1))
        np.testing.assert_array_equal(np.fromfile(act / "correct_classes.buffer", np.int32), lab)
        np.testing.assert_array_equal(np.fromfile(act / "softmax.buffer", np.float32), pred.reshape(-1))
        np.testing.assert_array_equal(np.fromfile(act / "max_inds.buffer", np.int32), t.activation("max_inds", dtype=np.int32))
        for fname, field in [("init_conv_applied", "init_conv_applied"), ("init_conv_activated", "init_conv_activated"),
                             ("init_convblock_input", "init_convblock_input"), ("final_avg_pool", "final_conv_output_pooled"),
                             ("fc_output", "linear_output")]:
            if field == "init_conv_activated" and not keep_all:
                # the stem's BatchNorm + ReLU + max pool are one kernel by default: the activated tensor is not materialised (and not dumped)
                assert t.activation(field) is None and not (act / (fname + ".buffer")).exists()
                continue
            np.testing.assert_array_equal(np.fromfile(act / (fname + ".buffer"), np.float32), t.activation(field))
        np.testing.assert_array_equal(np.fromfile(act / "batch_norms" / "init" / "means.buffer", np.float32), t.activation("norm_init_conv.means"))
        blk = {"reduction_applied": "post_reduced", "reduction_activated": "post_reduced_activated", "spatial_applied": "post_spatial",
               "spatial_activated": "post_spatial_activated", "expanded_applied": "post_expanded", "output_activated": "output_activated"}
        for bi in range(cfg["n_blocks"]):
            d = act / "conv_blocks" / ("%02d" % bi)
            for fname, field in blk.items():
                np.testing.assert_array_equal(np.fromfile(d / (fname + ".buffer"), np.float32), t.activation("b%d.%s" % (bi, field)))
            for sub in ("reduced", "spatial", "expanded"):
                assert (act / "batch_norms" / ("%02d" % bi) / sub / "vars.buffer").stat().st_size == 4 * {
                    "reduced": 64 * 2 ** sum(cfg["reductions"][:bi + 1]), "spatial": 64 * 2 ** sum(cfg["reductions"][:bi + 1]),
                    "expanded": 256 * 2 ** sum(cfg["reductions"][:bi + 1])}[sub]
            # pre-ReLU sums and BN outputs exist only when the trainer materialises them (keep-all), as in resnet_clean.h
            assert (d / "combined_output.buffer").exists() == keep_all
            assert (d / "expanded_post_norm.buffer").exists() == keep_all
        assert (act / "conv_blocks" / "00" / "transformed_residual.buffer").exists()           # block 0 projects 64 -> 256
        assert not (act / "conv_blocks" / "02" / "transformed_residual.buffer").exists()       # identity shortcut
        assert (root / "activation_derivs" / "conv_blocks" / "01" / "expanded_applied.buffer").exists()
        assert (root / "activation_derivs" / "softmax.buffer").stat().st_size == 4 * cfg["batch"] * cfg["output"]
        # ---- text files: line order of reference resnet.cu:2682-2753
        meta = (root / "trainer_metadata.txt").read_text().split("\n")
        assert int(meta[0]) == cfg["batch"] and int(meta[1]) == 3 * cfg["input_dim"] ** 2 and int(meta[2]) == cfg["input_dim"]
        assert abs(float(meta[4]) - cfg["lr"]) < 1e-6
        ck = (root / "trainer_checkpoint.txt").read_text().split()
        assert len(ck) == 6 and abs(float(ck[2]) - 0.9 ** 2) < 1e-6 and abs(float(ck[3]) - 0.999 ** 2) < 1e-6
        # ---- restore into a fresh trainer and take the same next step: bit-identical parameters
        t.update()
        t.set_batch(img, lab)
        pred_next = t.forward()
        t2 = make(cfg, dtype, keep_all)
        t2.restore(7, "run_a")
        tr2 = t2.t.contents
        assert tr2.init_loaded == 1 and abs(tr2.cur_mean_decay - 0.9 ** 2) < 1e-6
        for which in (0, 2, 3):
            for a, b in zip(t2.get_params(which), [np.fromfile(root / ["model_params", "gradients", "means", "vars"][which] / ("%03d.buffer" % i), np.float32)
                                                   for i in range(t2.n_locations)]):
                np.testing.assert_array_equal(a, b)
        t2.set_batch(img, lab)
        t2.forward()
        t2.backward()
        t2.update()
        for a, b in zip(t2.get_params(0), t.get_params(0)):
            np.testing.assert_array_equal(a, b)
        t2.set_batch(img, lab)
        np.testing.assert_array_equal(t2.forward(), pred_next)
        t.close()
        t2.close()
    finally:
        os.environ.pop("RESNET_B200_DUMP_ROOT", None)


def test_restore_missing_dump_reports_error(tmp_path):
    from resnet_b200 import api
    os.environ["RESNET_B200_DUMP_ROOT"] = str(tmp_path)
    try:
        t = make(G.MINI, "tf32", False)
        with pytest.raises(RuntimeError, match="cannot open"):
            t.restore(3, "nowhere")
        api.L().resnet_b200_clear_error()
        t.close()
    finally:
        os.environ.pop("RESNET_B200_DUMP_ROOT", None)


def test_on_device_epoch_bookkeeping():
    """loss / accuracy accumulate on the device across steps with no per-step synchronisation or pred_cpu copy;
    the sums equal the host-side definitions of the reference's loop (reference: resnet.cu:3363-3383) applied to pred."""
    cfg = G.MINI
    shapes = O.param_shapes(cfg["input_dim"], cfg["n_blocks"], cfg["reductions"], output=cfg["output"])
    t = make(cfg, "tf32", False)
    t.set_params(G.mini_weights(shapes))
    t.set_pred_copy(False)
    want_loss, want_wrong = 0.0, 0
    for step in range(3):
        img, lab = G.mini_batch(cfg, seed=21 + step)
        t.set_batch(img, lab)
        t.forward_async()
        pred = t.fetch_pred()               # on demand only: the loop itself does not need it
        p_lab = pred[np.arange(cfg["batch"]), lab]
        want_loss += float(-np.log(p_lab).sum())
        other = pred.copy()
        other[np.arange(cfg["batch"]), lab] = -1
        want_wrong += int((other.max(1) >= p_lab).sum())   # ties are wrong (reference: resnet.cu:3376)
        t.backward()
        t.update()
    loss, wrong, images = t.epoch_stats(reset=True)
    assert images == 3 * cfg["batch"] and wrong == want_wrong
    assert abs(loss - want_loss) < 1e-4 * abs(want_loss)
    assert t.epoch_stats() == (0.0, 0, 0)
    t.set_pred_copy(True)
    t.close()


def test_prefetch_commit_matches_blocking_copy():
    """resnet_b200_prefetch_batch / commit_batch (double-buffered H2D on a copy stream) deliver exactly the bytes a blocking copy
    does, including when the next batch is prefetched while the current one is still being consumed."""
    import ctypes as C
    from resnet_b200 import api
    cfg = G.MINI
    shapes = O.param_shapes(cfg["input_dim"], cfg["n_blocks"], cfg["reductions"], output=cfg["output"])
    t = make(cfg, "tf32", False)
    t.set_params(G.mini_weights(shapes))
    L = api.L()
    batches = [G.mini_batch(cfg, seed=30 + i) for i in range(3)]
    want = []
    for img, lab in batches:
        t.set_batch(img, lab)
        want.append(t.forward())
    pinned = []
    for img, lab in batches:
        hi, hl = L.resnet_b200_malloc_host(img.nbytes), L.resnet_b200_malloc_host(lab.nbytes)
        C.memmove(hi, img.ctypes.data, img.nbytes)
        C.memmove(hl, lab.ctypes.data, lab.nbytes)
        pinned.append((hi, hl))
    L.resnet_b200_prefetch_batch(t.t, *pinned[0])
    for i in range(3):
        L.resnet_b200_commit_batch(t.t)
        if i + 1 < 3:
            L.resnet_b200_prefetch_batch(t.t, *pinned[i + 1])   # overlaps with this step's forward
        np.testing.assert_array_equal(t.forward(), want[i])
    api.check()
    for hi, hl in pinned:
        L.resnet_b200_free_host(hi)
        L.resnet_b200_free_host(hl)
    t.close()
