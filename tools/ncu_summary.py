"""usage: tools/ncu_summary.py <report.ncu-rep> <out.json>  — per-kernel key metrics of an `ncu --set full` report."""
import csv, json, subprocess, sys, collections
rep, out = sys.argv[1], sys.argv[2]
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers", "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct"]
idx = {h: i for i, h in enumerate(hdr)}
res = collections.OrderedDict()
for r in data:
    name = r[idx["Kernel Name"]].split("(")[0]
    res.setdefault(name, [])
    if len(res[name]) < 3:
        res[name].append({k: f"{r[idx[k]]} {units[idx[k]]}".strip() for k in keys if k in idx})
json.dump(res, open(out, "w"), indent=1)
print(out, {k: len(v) for k, v in res.items()})
