"""Turn an `ncu --set full` report into profiles/<name>.json: per kernel the median over its captured launches of the metrics
the roofline blocks of bench.py and the summaries under profiles/ quote.

    python tools/ncu_metrics.py gpurun_out/r2_prof.ncu-rep profiles/r2_ncu_metrics.json
"""
import csv
import json
import statistics
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h, units = rows[0], rows[1]
col = {c: i for i, c in enumerate(h)}
M = {
    "gpu_time_us": ("gpu__time_duration.sum", 1e3),            # ms -> us (unit column is checked below)
    "dram_bytes_read": ("dram__bytes_read.sum", None),
    "dram_bytes_write": ("dram__bytes_write.sum", None),
    "sm_cycles_active_avg": ("sm__cycles_active.avg", 1),
    "sm_cycles_elapsed_max": ("sm__cycles_elapsed.max", 1),
    "issue_slots_pct_of_active": ("sm__inst_issued.avg.pct_of_peak_sustained_active", 1),
    "pipe_alu_pct_of_active": ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", 1),
    "pipe_fma_pct_of_active": ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", 1),
    "pipe_tensor_pct_of_active": ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 1),
    "sm_throughput_pct": ("sm__throughput.avg.pct_of_peak_sustained_elapsed", 1),
    "warp_instructions": ("smsp__inst_executed.sum", 1),
    "registers_per_thread": ("launch__registers_per_thread", 1),
    "grid": ("launch__grid_size", 1), "block": ("launch__block_size", 1), "cluster": ("launch__cluster_size", 1),
}
SCALE = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ms": 1e3, "us": 1, "ns": 1e-3, "s": 1e6}
per = {}
for r in rows[2:]:
    name = r[col["Kernel Name"]].split("(")[0].split("::")[-1].replace("void ", "").strip()
    name = name.split("<")[0]
    d = per.setdefault(name, {k: [] for k in M})
    d.setdefault("_launches", []).append(r[col["Kernel Name"]])
    for k, (m, _) in M.items():
        v = float(r[col[m]].replace(",", ""))
        u = units[col[m]]
        if k.startswith("dram") or k == "gpu_time_us":
            v *= SCALE[u]
        d[k].append(v)
res = {}
for name, d in per.items():
    launches = d.pop("_launches")
    # the auction kernel appears once per setting: keep the launches apart
    if name == "emd_auction_kernel" and len(launches) == 2:
        for tag, i in (("emd_auction_kernel", 0), ("emd_auction_kernel_train_setting", 1)):
            res[tag] = {k: v[i] for k, v in d.items()}
            res[tag]["dram_bytes_per_launch"] = res[tag]["dram_bytes_read"] + res[tag]["dram_bytes_write"]
        continue
    res[name] = {k: statistics.median(v) for k, v in d.items()}
    res[name]["launches_captured"] = len(launches)
    res[name]["dram_bytes_per_launch"] = res[name]["dram_bytes_read"] + res[name]["dram_bytes_write"]
res["_source"] = f"ncu --set full --clock-control none ({rep}); python tools/prof_target.py --emd-train; median over the captured launches"
json.dump(res, open(out, "w"), indent=1)
print(json.dumps(res, indent=1)[:1500])
