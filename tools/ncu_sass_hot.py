"""Top SASS instructions of an `ncu --page source --csv` dump by warp-stall samples, with the dominant stall reasons.
    ncu -i X.ncu-rep --page source --csv > src.csv ; python tools/ncu_sass_hot.py src.csv [N]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = rows[1]
col = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
data = []
for idx, r in enumerate(rows[2:]):
    if len(r) < len(hdr):
        continue
    s = int(r[col["# Samples"]] or 0)
    data.append((s, idx, r))
total = sum(d[0] for d in data)
print("kernel:", rows[0][1], " total samples:", total)
agg = {h: sum(int(d[2][col[h]] or 0) for d in data) for h in stall_cols}
print("stall reasons:", ", ".join("%s %.1f%%" % (h[6:], 100.0 * v / max(total, 1)) for h, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
for s, idx, r in sorted(data, key=lambda d: -d[0])[:topn]:
    reasons = sorted(((int(r[col[h]] or 0), h[6:]) for h in stall_cols), reverse=True)[:2]
    print("%6d %5.1f%%  #%-5d %-70s %s" % (s, 100.0 * s / max(total, 1), idx, r[col["Source"]].strip()[:70], " ".join("%s:%d" % (n, v) for v, n in reasons if v)))
