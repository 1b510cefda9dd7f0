"""Which convolution launches are bound by the MMA issue rate of one thread, and what would 2 / 4 issuing threads buy?
Joins the launch trace (plan shapes) with the ncu launch list of one step (duration, tensor-pipe active %, DRAM bytes) and applies
the measured issue law of profiles/r01_mma_rate.txt: clocks per tcgen05.mma with 1 / 2 / 4 issuing threads =
115.4 / 62.5 / 56.8 (N <= 64), 115.4 / 86.5 / 76.8 (N <= 128), 171 / 150.5 / 140.8 (N > 128; linear in between for the merged wgrad widths).
    python tools/issue_bound_model.py profiles/r01_ncu_all_kernels_tf32_final2.csv profiles/r01_launch_trace_tf32_final2.log"""
import csv
import re
import sys
from collections import OrderedDict

SM = 148
LAW = {64: (115.4, 62.5, 56.8), 128: (115.4, 86.5, 76.8), 256: (171.0, 150.5, 140.8)}


def law(n, i):
    if n <= 64:
        return LAW[64][i]
    if n <= 128:
        return LAW[128][i]
    a, b = LAW[128][i], LAW[256][i]
    return a + (b - a) * (min(n, 256) - 128) / 128.0


def plans(trace):
    out = []
    for line in open(trace):
        if not line.startswith("[tc_run]"):
            continue
        what = line[len("[tc_run] "):].split(" | ")[0]
        kv = dict(re.findall(r"(\w+)=([\w.+-]+)", line))
        kind, shape = what.split(" ", 1)
        k = int(shape.split("x")[0])
        d = dict(what=what, kind=kind, k=k, wgrad=" wgrad " in line)
        for key in ("m_tiles", "n_tiles", "BN", "kchunks", "k_boxes", "co_tiles", "m_pair", "ci_tiles", "taps", "tpt"):
            if key in kv:
                d[key] = int(kv[key])
        if d["wgrad"]:
            groups = -(-d["taps"] // d["tpt"])
            d["mmas"] = groups * (d["co_tiles"] // d["m_pair"]) * d["ci_tiles"] * d["k_boxes"] * d["m_pair"] * 4 / SM
            d["N"] = min(d["tpt"], d["taps"]) * d["BN"] if d["taps"] > 1 else d["BN"]
            d["N"] = d["taps"] * d["BN"] / groups   # average merged width
        else:
            taps = 7 if k == 7 else k * k
            d["mmas"] = d["m_tiles"] * d["n_tiles"] * taps * d["kchunks"] * 4 / SM
            d["N"] = d["BN"]
        out.append(d)
    return out


def launches(ncu_csv):
    rows = list(csv.reader(open(ncu_csv)))
    hdr, out = None, OrderedDict()
    for r in rows:
        if "ID" in r and "Metric Name" in r:
            hdr = r
            continue
        if hdr is None or len(r) != len(hdr):
            continue
        d = dict(zip(hdr, r))
        if "igemm" not in d["Kernel Name"]:
            continue
        out.setdefault(int(d["ID"]), {})[d["Metric Name"]] = float(d["Metric Value"].replace(",", ""))
    return list(out.values())


def main(ncu_csv, trace):
    P, L = plans(trace), launches(ncu_csv)
    n = min(len(P), len(L))
    P, L = P[-n:], L[-n:]          # the last step of both lists
    steps = 2 if len(P) % 2 == 0 and P[: len(P) // 2] == P[len(P) // 2:] else 1
    if steps == 2:
        P, L = P[len(P) // 2:], L[len(L) // 2:]
    agg = OrderedDict()
    for p, l in zip(P, L):
        t = l["gpu__time_duration.sum"] / 1000.0                      # us
        act = l["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"] / 100.0
        dram = (l["dram__bytes_read.sum"] + l["dram__bytes_write.sum"])
        floor_clk = 128.0 * p["N"] / 256.0
        cpm = floor_clk / max(act, 1e-6)                              # clocks per MMA this launch really took (clock independent)
        t_hbm = dram / 6539e9 * 1e6
        bound = law(p["N"], 0) / cpm                                  # 1.0 = exactly at the one-thread issue floor
        pred = []
        for i in (1, 2):
            t_issue = t * law(p["N"], i) / cpm if bound > 0.8 else t  # only issue-bound launches move
            pred.append(max(t_issue, min(t, t_hbm / 0.9)))
        key = p["what"]
        a = agg.setdefault(key, dict(n=0, t=0.0, t2=0.0, t4=0.0, N=p["N"], cpm=0.0, bound=0.0, hbm=0.0))
        a["n"] += 1; a["t"] += t; a["t2"] += pred[0]; a["t4"] += pred[1]; a["cpm"] += cpm; a["bound"] += bound; a["hbm"] += t_hbm
    print("launch class                          n   N_mma  clk/MMA  issue-floor/actual  HBM-time/actual   us now -> 2 issuers -> 4 issuers")
    T = T2 = T4 = 0.0
    for k, a in agg.items():
        n = a["n"]
        print("%-36s %2d  %6.0f  %7.0f  %12.2f  %15.2f   %8.1f -> %8.1f -> %8.1f" % (k, n, a["N"], a["cpm"] / n, a["bound"] / n, a["hbm"] / a["t"], a["t"], a["t2"], a["t4"]))
        T += a["t"]; T2 += a["t2"]; T4 += a["t4"]
    print("all convolution launches of the step: %.2f ms now -> %.2f ms (2 issuers) -> %.2f ms (4 issuers), if every launch within 20 %% of the issue floor"
          " follows the microbenchmark law and nothing goes below 0.9 x its HBM time" % (T / 1e3, T2 / 1e3, T4 / 1e3))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
