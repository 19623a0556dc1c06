"""Summarise an `ncu --page source --csv` dump: samples per contiguous SASS region and top stall reasons.

usage: python tools/ncu_source_summary.py <source.csv> [region_size]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[ix["# Samples"]] or 0) for r in data)
print("total samples", tot, "instructions", len(data))
# regions split at branch targets is hard; use loops: print a running table of every instruction with >0.4% samples
acc = 0
print("%6s %7s %6s %9s  %s" % ("line", "samples", "pct", "exec", "sass / top stalls"))
for n, r in enumerate(data):
    s = int(r[ix["# Samples"]] or 0)
    acc += s
    if s >= tot * float(sys.argv[2] if len(sys.argv) > 2 else 0.004):
        st = sorted(((int(r[ix[c]] or 0), c) for c in stall_cols), reverse=True)[:3]
        print("%6d %7d %5.1f%% %9s  %-60s %s" % (n, s, 100.0 * s / tot, r[ix["Instructions Executed"]], r[ix["Source"]][:60],
                                             " ".join("%s=%d" % (c[6:], v) for v, c in st if v)))
# overall stall mix
mix = {c: sum(int(r[ix[c]] or 0) for r in data) for c in stall_cols}
print("stall mix:", " ".join("%s=%.1f%%" % (c[6:], 100.0 * v / tot) for c, v in sorted(mix.items(), key=lambda kv: -kv[1]) if v > tot * 0.005))
# cumulative samples by 100-instruction windows
w = 100
for a in range(0, len(data), w):
    s = sum(int(r[ix["# Samples"]] or 0) for r in data[a:a + w])
    e = sum(int(r[ix["Instructions Executed"]] or 0) for r in data[a:a + w])
    print("instr %5d-%5d: samples %5.1f%%  executed %5.1f%%" % (a, a + w - 1, 100.0 * s / tot, 100.0 * e / max(1, sum(int(r[ix["Instructions Executed"]] or 0) for r in data))))
