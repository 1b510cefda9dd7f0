"""bench.py -- ResNet-50 training throughput of the B200-native hot path (BASELINE.json metric), one JSON line.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config c2|c3|c4|c5] [--impl ours|reference|reference_cached] [--batch B]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

Workload (BASELINE.json configs[1] = c2): the reference's ResNet-50 variant (3x3/2 projection shortcuts, 47.6 M parameters,
39.09 GFLOP / image / training step), full training step = batch in -> forward_pass -> backwards_pass ->
update_parameters (Adam), batch 256 per GPU, 224x224 synthetic images, fp32 storage, TF32 tensor-core convolutions
with fp32 accumulation.  N > 1: batch-sharded data parallel, one process per GPU, NCCL gradient allreduce (weak scaling).

  value    images/s with the batch already resident in HBM (device->device restore of cur_batch each step, because the
           reference's update_parameters zeroes it); CUDA events on the trainer's stream; max over ranks.
  e2e      the same step through the public C API with HOST buffers: pinned host -> device copy of images + labels
           (double-buffered on a copy stream: resnet_b200_prefetch_batch / commit_batch, so batch k+1 crosses PCIe while step
           k computes) and the device -> host read of pred_cpu inside the timed region; K copies complete inside K steps.
  roofline dominant kernel family, timed per launch with CUDA events on the launching stream in an instrumented pass
           of the same step right after the timed region (the event pairs would perturb the headline number).
  cpu_baseline  the host-core C oracle (port of the reference's kernels; the reference has no CPU path) on a bounded
           sample of the same workload.
  extra    (default c2 run only, --no-extra switches it off) the other BASELINE configurations measured the same way right after
           the headline, a few steps each: c4 (bf16 training, the data-parallel scaling configuration), c5 (ResNet-152 bf16, batch
           128 per GPU) at every N, c3 (forward only, bf16, batch 1024) at N = 1.  Each entry: value, ms_per_step, e2e, roofline_all.
  --impl reference  the reference's own resnet_cudnn_fast.cu (oracle/_ref/libref_fast.so, unmodified sources, its own
           forward_pass/backwards_pass/update_parameters) on the same config on this GPU (--config c3: its forward_pass only, batch
           1024); falls back to the oracle port on host cores when that library cannot run.
  --impl reference_cached  the same translation unit compiled with a caching cudaMalloc / cudaFree shim (oracle/ref_harness.cu,
           -DREF_CACHING_ALLOC): the stock build allocates and frees a workspace around every convolution call, so `reference`
           measures allocator stalls as much as cuDNN 9.10's kernels.  A LABELLED extra arm -- never the driver's baseline.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

R50_REDUCTIONS = [1 if i in (3, 7, 13) else 0 for i in range(16)]
R152_REDUCTIONS = [1 if i in (3, 11, 47) else 0 for i in range(50)]
FLOP_PER_IMAGE_STEP = 39.09e9  # fprop + dgrad + wgrad, no stem dgrad
METRIC = "ResNet-50 train img/s"
# BASELINE.json configs: c2 is the bench line (configs[1]); c3-c5 are the bf16 configurations, measured with --config
CONFIGS = {
    "c2": dict(name="ResNet-50", red=R50_REDUCTIONS, batch=256, dtype="tf32", fwd_only=False, gflop=39.09),
    "c3": dict(name="ResNet-50", red=R50_REDUCTIONS, batch=1024, dtype="bf16", fwd_only=True, gflop=13.111),
    "c4": dict(name="ResNet-50", red=R50_REDUCTIONS, batch=256, dtype="bf16", fwd_only=False, gflop=39.09),
    "c5": dict(name="ResNet-152", red=R152_REDUCTIONS, batch=128, dtype="bf16", fwd_only=False, gflop=83.64),
}


def peaks():
    p = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "src": "fallback"}
    f = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(f):
        try:
            p.update(json.load(open(f)))
            p["src"] = "measured"
        except Exception:
            pass
    return p


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def init_dist(world):
    if world <= 1:
        return None
    import torch.distributed as dist
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29511")
    dist.init_process_group(backend="gloo")  # plumbing only: rendezvous, barrier, max-reduce of timings, NCCL id broadcast
    return dist


def barrier(dist):
    if dist is not None:
        dist.barrier()


def reduce_max(dist, x):
    if dist is None:
        return x
    import torch
    t = torch.tensor([x], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def broadcast_bytes(dist, payload, n, src=0):
    """rank `src`'s `payload` (n bytes) on every rank -- carries the NCCL unique id from rank 0 to the others"""
    if dist is None:
        return bytes(payload)
    import torch
    t = torch.tensor(list(bytes(payload)) if dist.get_rank() == src else [0] * n, dtype=torch.uint8)
    dist.broadcast(t, src=src)
    return bytes(t.tolist())


def reduce_sum(dist, x):
    if dist is None:
        return x
    import torch
    t = torch.tensor([x], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def metric_of(cfg_name):
    cfg = CONFIGS[cfg_name]
    return METRIC if cfg_name in ("c2", "c4") else ("ResNet-50 forward img/s" if cfg["fwd_only"] else "ResNet-152 train img/s")


def config_dict(cfg_name, batch, world):
    """`config` of the JSON line: a description of the WORKLOAD only, identical for the product arm and the reference arms (arm-specific
    detail goes to `impl_detail`)."""
    cfg = CONFIGS[cfg_name]
    what = "forward_pass only (batch-statistics BatchNorm: the reference has no inference mode)" if cfg["fwd_only"] else \
        "full training step (forward_pass + backwards_pass + Adam update_parameters)"
    return {"workload": "%s (reference variant: 3x3/2 projection shortcuts, %.2f GFLOP/img/step) %s, batch %d per GPU, 224x224 synthetic images"
                        % (cfg["name"], cfg["gflop"], what, batch),
            "baseline_config": cfg_name, "global_batch": batch * world, "per_gpu_batch": batch, "parallelism": "dp%d" % world,
            "l2": "inputs larger than L2 (tens of GB touched per step), no flush needed"}


def pick_dominant(rl_all):
    """The kernel family the headline `roofline` object describes: the one with the largest share of the step (the two halves of the
    tcgen05 fprop + dgrad family -- entries named "igemm_kmajor_kernel, ..." -- are listed for the reader and do not compete).
    The tcgen05 fprop + dgrad family and the BatchNorm / elementwise family are within 1-2 % of each other in the ResNet-50 step
    (about 16.5 ms each): the headline stays on the tensor-core kernel -- the one earlier rounds and reviews quote -- unless another
    family leads it by more than 5 %.  Every family is in roofline_all either way."""
    cands = [r for r in rl_all if not r["kernel"].startswith("igemm_kmajor_kernel, ")]
    dom = max(cands, key=lambda r: r["ms_per_step"])
    km = [r for r in cands if r["kernel"].startswith("igemm_kmajor_kernel")]
    if km and km[0]["ms_per_step"] >= 0.95 * dom["ms_per_step"]:
        dom = km[0]
    return dom


# ------------------------------------------------------------------------------------------------ CPU baseline (oracle port)
def cpu_baseline(batch=4):
    from oracle import oracle as O
    from resnet_b200.synth import synthetic_batch
    net = O.OracleNet(224, 16, R50_REDUCTIONS, batch=batch)
    net.init_like_reference(0)
    img, lab = synthetic_batch(batch, 224, seed=1234)
    t0 = time.time()
    net.forward(img, lab)
    net.backward()
    net.update()
    dt = time.time() - t0
    return {"value": batch / dt, "unit": "img/s", "cores": O.num_threads(), "kind": "port",
            "sample": "1 full training step (forward, backward, Adam) of the same ResNet-50 at batch %d, oracle/ops.c with OpenMP, %.1f s" % (batch, dt)}


# ------------------------------------------------------------------------------------------------ reference arms
def run_reference(args, rank, world, cached=False):
    """the reference's own cuDNN build on this GPU (rank 0 only; the reference is single-GPU).  cached: the caching-allocator build."""
    if rank != 0:
        return
    cfg = CONFIGS[args.config]
    fwd_only = cfg["fwd_only"]
    variant = "fast_cached" if cached else "fast"
    out = {"impl": "reference_cached" if cached else "reference", "metric": metric_of(args.config), "unit": "img/s", "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "data": "synthetic",
           "config": config_dict(args.config, args.batch, max(1, args.gpus))}
    from oracle import ref as R
    from resnet_b200.synth import synthetic_batch
    n_blocks = len(cfg["red"])
    done = False
    if n_blocks != 16:
        out["reference_gpu_error"] = "the reference's init_resnet hard-codes ResNet-50's four projection blocks; no ResNet-152 baseline exists"
    elif R.available(variant):
        try:
            img, lab = synthetic_batch(args.batch, 224, seed=1234)
            # the reference's cuDNN build, unmodified: its own init_*, forward_pass, backwards_pass, update_parameters
            r = R.Ref(variant).create(224, 16, cfg["red"], args.batch, output=1000, lr=1e-3, seed=1234)
            r.set_batch(img, lab)
            ms = r.time_steps(args.warmup, args.steps, e2e=False, forward_only=fwd_only)
            ms_e2e = r.time_steps(1, max(2, args.steps // 2), e2e=True, forward_only=fwd_only)
            v = args.batch / (ms / 1e3)
            detail = ("the reference's resnet_cudnn_fast.cu (cuDNN 9.10, fp32 NCHW, TENSOR_OP_MATH_ALLOW_CONVERSION = TF32 tensor-core math), compiled "
                      "-O3 --use_fast_math sm_100a, its own %s; single-GPU: runs on GPU 0 for every --gpus" %
                      ("forward_pass" if fwd_only else "forward_pass / backwards_pass / update_parameters"))
            if cached:
                detail += "; cudaMalloc / cudaFree inside the step served by a caching shim (-DREF_CACHING_ALLOC): NOT the stock reference"
            out.update({"value": v, "ms_per_step": ms, "dtype": "tf32", "gpu_launches": None, "impl_detail": detail, "n_gpus_used": 1,
                        "cpu_baseline": {"value": v, "unit": "img/s", "cores": 0, "kind": "reference",
                                         "sample": "GPU run of oracle/_ref/%s (the reference has no CPU path); cuda error: %s" % (R.VARIANTS[variant], r.cuda_error())},
                        "e2e": {"value": args.batch / (ms_e2e / 1e3), "unit": "img/s", "h2d_bytes_per_step": int(img.nbytes + lab.nbytes),
                                "d2h_bytes_per_step": int(args.batch * 1000 * 4)}})
            done = np.isfinite(v) and v > 0
        except Exception as e:  # noqa: BLE001
            out["reference_gpu_error"] = repr(e)
    if not done:
        cb = cpu_baseline(batch=4)
        out.update({"value": cb["value"], "ms_per_step": 1e3 * 4 / cb["value"], "dtype": "f32", "gpu_launches": 0, "cpu_baseline": cb,
                    "impl_detail": "host-core port of the reference's kernels (oracle/ops.c), one ResNet-50 training step at batch 4",
                    "e2e": {"value": cb["value"], "unit": "img/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
    print(json.dumps(out), flush=True)


# ------------------------------------------------------------------------------------------------ our arm
def dp_join(L, api, t, dist, rank, world):
    """joins trainer `t` to the data-parallel world: NCCL id from rank 0 over the gloo group, then resnet_b200_dp_init"""
    idbuf = (C.c_char * 128)()
    if rank == 0:
        L.resnet_b200_dp_unique_id(idbuf)
    idbuf = (C.c_char * 128)(*broadcast_bytes(dist, bytes(idbuf), 128))
    # NCCL prints its version banner on stdout when NCCL_DEBUG=VERSION is set in the environment; stdout carries exactly one
    # JSON line (the driver parses it), so the communicator is created with fd 1 pointing at stderr
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    try:
        L.resnet_b200_dp_init(t.t, idbuf, rank, world, 0)
    finally:
        os.dup2(saved, 1)
        os.close(saved)
    api.check()


def measure(cfg_name, batch, steps, warmup, dist, rank, local_rank, world, with_clocks=True):
    """One configuration through the C API; returns the result dict on rank 0 (None elsewhere).  The trainer is destroyed on return."""
    from resnet_b200 import api
    from resnet_b200.synth import synthetic_batch
    L = api.L()
    cfg = CONFIGS[cfg_name]
    pk = peaks()
    N = batch
    t = api.Trainer(input_dim=224, n_blocks=len(cfg["red"]), reductions=cfg["red"], batch=N, output=1000, lr=1e-4, seed=1234, device=local_rank,
                    dtype=cfg["dtype"])
    assert t.uses_tensor_cores(), "bench must run the tcgen05 path"
    assert t.bf16 == (cfg["dtype"] == "bf16")
    fwd_only = cfg["fwd_only"]
    flop_per_image = cfg["gflop"] * 1e9
    if world > 1:
        dp_join(L, api, t, dist, rank, world)

    img, lab = synthetic_batch(N, 224, seed=1234 + 1000 * rank)
    host_img = L.resnet_b200_malloc_host(img.nbytes)
    host_lab = L.resnet_b200_malloc_host(lab.nbytes)
    C.memmove(host_img, img.ctypes.data, img.nbytes)
    C.memmove(host_lab, lab.ctypes.data, lab.nbytes)
    dev_img = api.DevBuf(img)
    dev_lab = api.DevBuf(lab)

    def step(e2e):
        if e2e:
            # this step's batch was prefetched (pinned host -> staging buffer on the copy stream) while the previous step computed;
            # commit it, then start the copy of the next one.  Every step still moves its own 154 MB across PCIe inside the timed region.
            L.resnet_b200_commit_batch(t.t)
            L.resnet_b200_prefetch_batch(t.t, host_img, host_lab)
        else:
            L.resnet_b200_stage_batch_device(t.t, dev_img.ptr, dev_lab.ptr)
        L.forward_pass(t.t)                                   # returns with pred_cpu valid (D2H inside)
        _ = t.t.contents.forward_buffer.contents.pred_cpu[0]  # the host reads the prediction, as the reference's loop does
        if not fwd_only:
            L.backwards_pass(t.t)
            L.update_parameters(t.t)

    for _ in range(max(3, warmup)):
        step(False)
    t.sync()
    api.check()

    # ---- timed region: value (batch resident in HBM)
    barrier(dist)
    t.sync()
    sampler = ClockSampler(local_rank) if (rank == 0 and with_clocks) else None
    launches0 = L.resnet_b200_launch_count()
    L.resnet_b200_timer_begin(t.t)
    for _ in range(steps):
        step(False)
    ms = L.resnet_b200_timer_end_ms(t.t)
    t.sync()
    launches = L.resnet_b200_launch_count() - launches0
    barrier(dist)
    clocks = sampler.stop() if sampler else None
    ms_max = reduce_max(dist, ms)
    api.check()

    # ---- timed region: e2e (host buffers, copies inside)
    barrier(dist)
    L.resnet_b200_prefetch_batch(t.t, host_img, host_lab)   # pipeline fill: step 0's batch
    t.sync()
    L.resnet_b200_timer_begin(t.t)
    for _ in range(steps):
        step(True)                                          # commit batch k, start the copy of batch k+1, run the step
    L.resnet_b200_commit_batch(t.t)                         # drain: K host->device copies have completed inside the timed region
    ms_e2e = L.resnet_b200_timer_end_ms(t.t)
    t.sync()
    ms_e2e_max = reduce_max(dist, ms_e2e)
    loss, nwrong = t.loss_accuracy()

    # ---- instrumented pass: per-family CUDA-event timing
    L.resnet_b200_profile(1)
    prof_steps = 2
    for _ in range(prof_steps):
        step(False)
    t.sync()
    fam = {}
    for f, name in ((0, "igemm_kmajor_kernel, 3x3 + stem fprop/dgrad"), (1, "igemm_mnmajor_kernel (tcgen05 wgrad + split-K reduce)"),
                    (2, "BatchNorm/elementwise"), (3, "simt_conv_kernel (stem 7x7, fp32)"), (5, "igemm_kmajor_kernel, 1x1 fprop/dgrad")):
        tms, n, w, w2 = C.c_double(), C.c_longlong(), C.c_double(), C.c_double()
        L.resnet_b200_profile_read2(f, C.byref(tms), C.byref(n), C.byref(w), C.byref(w2))
        fam[f] = {"name": name, "ms": tms.value, "launches": n.value, "work": w.value, "work2": w2.value}
    # the whole fprop + dgrad kernel family (what round 1 reported as one line): 3x3 / stem launches + 1x1 launches
    fam[6] = {"name": "igemm_kmajor_kernel (tcgen05 fprop+dgrad)", "ms": fam[0]["ms"] + fam[5]["ms"], "launches": fam[0]["launches"] + fam[5]["launches"],
              "work": fam[0]["work"] + fam[5]["work"], "work2": 0.0}
    L.resnet_b200_profile(0)
    api.check()

    out = None
    if rank == 0:
        tensor_peak = pk["bf16_tflops_sustained"] / (1.0 if t.bf16 else 2.0)  # tensor peak of the MMA kind in use
        rl_all = []
        for f in (6, 0, 1, 3):
            if fam[f]["launches"]:
                ach = fam[f]["work"] / (fam[f]["ms"] * 1e-3) / 1e12
                peak = tensor_peak if f != 3 else 75.0
                rl_all.append({"kernel": fam[f]["name"], "bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
                               "ms_per_step": fam[f]["ms"] / prof_steps, "launches_per_step": fam[f]["launches"] / prof_steps,
                               "flops_per_launch": fam[f]["work"] / fam[f]["launches"]})
        if fam[2]["launches"]:
            ach = fam[2]["work"] / (fam[2]["ms"] * 1e-3) / 1e9
            rl_all.append({"kernel": fam[2]["name"], "bound": "hbm", "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": ach / pk["hbm_gbs"],
                           "ms_per_step": fam[2]["ms"] / prof_steps, "launches_per_step": fam[2]["launches"] / prof_steps,
                           "bytes_per_launch": fam[2]["work"] / fam[2]["launches"]})
        if fam[5]["launches"]:
            # the 1x1 convolutions move their input and output tensor once and have 4-16x fewer FLOPs per byte than the 3x3s: HBM-bound
            ach = fam[5]["work2"] / (fam[5]["ms"] * 1e-3) / 1e9
            rl_all.append({"kernel": fam[5]["name"], "bound": "hbm", "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": ach / pk["hbm_gbs"],
                           "ms_per_step": fam[5]["ms"] / prof_steps, "launches_per_step": fam[5]["launches"] / prof_steps,
                           "bytes_per_launch": fam[5]["work2"] / fam[5]["launches"],
                           "tflops": fam[5]["work"] / (fam[5]["ms"] * 1e-3) / 1e12})
        # dominant = the kernel family with the largest share of the step (the whole-family kmajor line competes as one entry; its two
        # halves are listed after it for the reader)
        dom = pick_dominant(rl_all)
        # DRAM traffic per launch of the dominant family from the committed ncu capture of one step of this configuration's dtype
        # (profiles/r0N_traffic_<dtype>.json, written by tools/ncu_summary.py --json; batch 256 ResNet-50 only; newest round first)
        traffic, traffic_src = None, None
        for rnd in ("r02", "r01"):
            tfile = os.path.join(ROOT, "profiles", "%s_traffic_%s.json" % (rnd, cfg["dtype"]))
            if traffic is None and os.path.exists(tfile) and cfg_name in ("c2", "c4") and N == 256:
                try:
                    tj = json.load(open(tfile))
                    key = "kmajor" if "kmajor" in dom["kernel"] else ("wgrad" if "mnmajor" in dom["kernel"] else "bn_eltwise")
                    traffic = tj[key]["dram_bytes_per_step"] / dom["launches_per_step"]
                    traffic_src = "profiles/%s_traffic_%s.json (ncu dram__bytes_read.sum + dram__bytes_write.sum of one step / launches of the family)" % (rnd, cfg["dtype"])
                except Exception:
                    traffic = None
        roofline = {"bound": dom["bound"], "achieved": dom["achieved"], "peak": dom["peak"], "unit": dom["unit"], "frac": dom["frac"],
                    "traffic": traffic, "traffic_source": traffic_src, "kernel": dom["kernel"],
                    "peak_source": ("MEASURED_PEAKS.json (%s): " % pk["src"]) + (("bf16_tflops_sustained (kernel timed inside a long step)" if t.bf16 else
                                                                                  "bf16_tflops_sustained / 2 (TF32 runs at half the bf16 rate; kernel timed inside a long step)")
                                                                                 if dom["bound"] == "tensor" else "hbm_gbs copy bandwidth"),
                    "how": "CUDA events around every launch of the family on the launching stream, %d instrumented steps after the timed region" % prof_steps}
        total_img = N * world * steps
        value = total_img / (ms_max * 1e-3)
        config = config_dict(cfg_name, N, world)
        out = {"metric": metric_of(cfg_name), "value": value, "unit": "img/s", "n_gpus": world, "steps": steps, "warmup": max(3, warmup),
               "ms_per_step": ms_max / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": cfg["dtype"], "data": "synthetic",
               "config": config,
               "impl_detail": "libresnet_b200.so through the reference's entry points (ctypes): %s; %.2f M parameters; step FLOPs %.4g; %.1f TFLOP/s per GPU"
                              % ("bf16 storage NHWC, bf16 tcgen05 convolutions, fp32 master weights / gradients / Adam" if t.bf16 else
                                 "fp32 storage NHWC, TF32 tcgen05 convolutions", sum(t.sizes) / 1e6, flop_per_image * N,
                                 flop_per_image * N / (ms_max / steps * 1e-3) / 1e12),
               "e2e": {"value": total_img / (ms_e2e_max * 1e-3), "unit": "img/s", "h2d_bytes_per_step": int(img.nbytes + lab.nbytes),
                       "d2h_bytes_per_step": int(N * 1000 * 4), "ms_per_step": ms_e2e_max / steps},
               "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "roofline_all": rl_all,
               "last_step": {"loss_per_image": loss / N, "n_wrong": nwrong}}
    barrier(dist)
    t.close()
    dev_img.free()
    dev_lab.free()
    L.resnet_b200_free_host(host_img)
    L.resnet_b200_free_host(host_lab)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=None, help="images per GPU (default: the config's)")
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS), help="BASELINE.json configuration (default c2 = configs[1], the bench line)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "reference_cached"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the c3 / c4 / c5 measurements that follow the default c2 headline")
    args = ap.parse_args()
    cfg = CONFIGS[args.config]
    default_run = args.config == "c2" and args.batch is None
    if args.batch is None:
        args.batch = cfg["batch"]
    rank, local_rank, world = dist_env()
    if args.impl != "ours":
        run_reference(args, rank, world, cached=(args.impl == "reference_cached"))
        return

    dist = init_dist(world)
    from resnet_b200 import api
    api.L().resnet_b200_set_device(local_rank)
    out = measure(args.config, args.batch, args.steps, args.warmup, dist, rank, local_rank, world)
    if default_run and not args.no_extra:
        # the other BASELINE configurations, a few steps each, so that the driver's N = 1 / 2 / 4 / 8 runs also record the bf16
        # data-parallel configuration (c4), ResNet-152 (c5) and, on one GPU, forward-only inference (c3)
        extra = {}
        for name in (("c4", "c5", "c3") if world == 1 else ("c4", "c5")):
            try:
                r = measure(name, CONFIGS[name]["batch"], min(args.steps, 8), 3, dist, rank, local_rank, world, with_clocks=False)
            except Exception as e:  # noqa: BLE001  (an extra must never take the headline down)
                r = {"error": repr(e)} if rank == 0 else None
            if rank == 0 and r is not None:
                keep = ("metric", "value", "unit", "ms_per_step", "dtype", "config", "e2e", "gpu_launches", "roofline_all", "last_step", "error")
                extra[name] = {k: r[k] for k in keep if k in r}
        if rank == 0:
            out["extra"] = extra
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_baseline(batch=4)
        print(json.dumps(out), flush=True)
    barrier(dist)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
