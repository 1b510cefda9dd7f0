"""Turns an `ncu --csv --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active...,dram__bytes_*` launch list
(long format: one row per launch and metric) into the per-launch / per-kernel tables committed under profiles/.

    python tools/ncu_summary.py launches.csv [--trace trace.log] [--skip N] > profiles/rNN_....txt

--trace: stderr of the same program run with RESNET_B200_TRACE=1 (one "[tc_run] <layer> | <plan> flops=F" line per tensor-core
launch, in launch order); igemm launches are then labelled with their layer and get achieved TFLOP/s."""
import argparse
import collections
import csv
import re

ap = argparse.ArgumentParser()
ap.add_argument("csv")
ap.add_argument("--trace")
ap.add_argument("--skip", type=int, default=0, help="ignore the first N launches (warm-up step)")
ap.add_argument("--last-step", action="store_true", help="start at the last pack_weights launch (the first kernel of forward_pass)")
ap.add_argument("--json", help="also write per-family DRAM traffic per launch (bench.py's roofline.traffic) to this file")
ap.add_argument("--hbm-peak", type=float, default=6539.2)
a = ap.parse_args()

rows = list(csv.reader(l for l in open(a.csv) if l.startswith('"')))
hdr = rows[0]
iid, ik, im, iv = hdr.index("ID"), hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
launch = collections.OrderedDict()
for r in rows[1:]:
    d = launch.setdefault(int(r[iid]), {"kernel": re.sub(r"\(.*", "", r[ik]).replace("rb::", "")})
    d[r[im]] = float(r[iv].replace(",", ""))
if a.last_step:
    order = sorted(launch)
    starts = [n for n, i in enumerate(order) if launch[i]["kernel"].endswith("pack_weights_kernel") or "pack_weights_kernel<" in launch[i]["kernel"]]
    starts = [n for n in starts if "stem" not in launch[order[n]]["kernel"]]
    a.skip = starts[-1] if starts else 0
ids = sorted(launch)[a.skip:]
trace = []
ktrace = collections.defaultdict(list)   # "[k] <kernel> <label>" lines, one queue per kernel name
if a.trace:
    for l in open(a.trace):
        m = re.match(r"\[tc_run\] (.*?) \| .* flops=([0-9.e+]+)", l)
        if m:
            trace.append((m.group(1), float(m.group(2))))
        m = re.match(r"\[k\] (\S+) (.*)", l)
        if m:
            ktrace[m.group(1)].append(m.group(2).strip())
kcount = collections.Counter()
for i in sorted(launch):
    for k in ktrace:
        if k in launch[i]["kernel"]:
            kcount[k] += 1
kpos = {k: len(ktrace[k]) - kcount[k] for k in ktrace}   # align queue tails with the captured launches
for i in sorted(launch)[:a.skip]:
    for k in ktrace:
        if k in launch[i]["kernel"]:
            kpos[k] += 1
n_tc_total = sum(1 for i in sorted(launch) if "igemm" in launch[i]["kernel"])
ti = len(trace) - n_tc_total if trace else 0   # the trace covers the whole run; align its tail with the captured launches
for i in sorted(launch)[:a.skip]:
    if "igemm" in launch[i]["kernel"]:
        ti += 1
print("columns: kernel | layer | duration us | tensor-pipe active %% | DRAM GB/s (read+write)/duration | %% of %.0f GB/s | TFLOP/s" % a.hbm_peak)
agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
for i in ids:
    d = launch[i]
    us = d.get("gpu__time_duration.sum", 0.0) / 1e3
    tp = d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0.0)
    gbs = (d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)) / max(us, 1e-9) / 1e3
    layer, tf = "", ""
    if "igemm" in d["kernel"] and trace and 0 <= ti < len(trace):
        layer, fl = trace[ti]
        tf = "%.0f" % (fl / (us * 1e-6) / 1e12)
        agg_key = d["kernel"] + " " + layer.split()[0]
        ti += 1
    else:
        if "igemm" in d["kernel"]:
            ti += 1
        agg_key = d["kernel"]
        for k in ktrace:
            if k in d["kernel"]:
                if 0 <= kpos[k] < len(ktrace[k]):
                    layer = ktrace[k][kpos[k]]
                kpos[k] += 1
    print("%-34s %-46s %9.1f %6.1f %8.1f %6.1f %6s" % (d["kernel"][:34], layer[:46], us, tp, gbs, 100 * gbs / a.hbm_peak, tf))
    g = agg[agg_key]
    g[0] += 1; g[1] += us; g[2] += tp * us; g[3] += gbs * us
tot = sum(v[1] for v in agg.values())
print("\nper kernel (time-weighted averages), total %.3f ms over %d launches" % (tot / 1e3, len(ids)))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-44s n=%4d total=%8.3f ms %5.1f%%  avg=%8.1f us  tensor %5.1f%%  DRAM %7.1f GB/s" % (k[:44], v[0], v[1] / 1e3, 100 * v[1] / tot, v[1] / v[0], v[2] / v[1], v[3] / v[1]))

if a.json:
    import json
    fam = {"kmajor": ["igemm_kmajor_kernel"], "wgrad": ["igemm_mnmajor_kernel", "wgrad_reduce"],
           "bn_eltwise": ["bn_apply_kernel", "bn_reduce_kernel", "bn_bwd_dx_kernel", "relu_bwd_kernel", "bn_finalize_kernel", "bn_bwd_finalize_kernel",
                          "bn_pool_fwd_kernel", "pool_bn_bwd_kernel"]}
    out = {"source": a.csv, "note": "dram__bytes_read.sum + dram__bytes_write.sum per launch, one step under ncu (cold caches, serialised)"}
    for f, pats in fam.items():
        n, by, us = 0, 0.0, 0.0
        for i in ids:
            d = launch[i]
            if any(p in d["kernel"] for p in pats):
                # the family's launch count follows bench.py's ProfScope granularity: reduces / finalizes ride inside their parent's scope
                main = any(p in d["kernel"] for p in pats[:1]) if f != "bn_eltwise" else ("finalize" not in d["kernel"] and "bn_reduce" not in d["kernel"] and
                                                                                                not ("pool_bn_bwd" in d["kernel"] and ", 0>" in d["kernel"]))
                n += 1 if main else 0
                by += d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
                us += d.get("gpu__time_duration.sum", 0.0) / 1e3
        out[f] = {"launches": n, "dram_bytes_per_launch": by / max(n, 1), "dram_bytes_per_step": by, "kernel_us_per_step": us}
    json.dump(out, open(a.json, "w"), indent=1)
