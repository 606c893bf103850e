#!/usr/bin/env python
"""tools/ncu_summarize.py <tag> <out-prefix> : turn the raw ncu outputs of tools/gpu_check.sh (gpurun_out/<tag>_launches.csv,
gpurun_out/<tag>_prof.ncu-rep) into the committed summaries profiles/<out-prefix>_launches.csv (time per kernel, shares)
and profiles/<out-prefix>_ncu_full.csv (key metrics of every captured launch of the full set)."""
import csv
import io
import re
import subprocess
import sys
from collections import OrderedDict

tag, outp = sys.argv[1], sys.argv[2]
CMD = (sys.argv[3] if len(sys.argv) > 3 else
       "python bench.py --steps 2 --warmup 3 --burn-in 1 --repeats 1 --no-extra --no-cpu-baseline --no-torch-baseline --no-profile")


RX = sys.argv[4] if len(sys.argv) > 4 else "fused_layers_kernel|final_edge_tc_kernel"


def short(name):
    name = re.sub(r"^void\s+", "", name)
    name = re.sub(r"gdss::", "", name)
    name = re.sub(r"\(.*$", "", name)
    return name.replace("(bool)", "")


rows = [l for l in open(f"gpurun_out/{tag}_launches.csv") if l.startswith('"')]
agg = OrderedDict()
for r in csv.DictReader(io.StringIO("".join(rows))):
    if r["Metric Name"] != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    if r["Metric Unit"] in ("nsecond", "ns"):
        v /= 1e3
    elif r["Metric Unit"] in ("msecond", "ms"):
        v *= 1e3
    k = short(r["Kernel Name"])
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
with open(f"profiles/{outp}_launches.csv", "w") as f:
    f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none : {CMD}\n")
    f.write("# per-launch times are cold-cache and serialised: compare SHARES with bench.py's kernel_ms shares\n")
    f.write("kernel,launches,total_us,share\n")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"\"{k}\",{a[0]},{a[1]:.1f},{a[1] / tot:.4f}\n")

METRICS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
           "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__inst_issued.avg.pct_of_peak_sustained_active",
           "smsp__warps_eligible.avg.per_cycle_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
           "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]
raw = subprocess.run(["ncu", "-i", f"gpurun_out/{tag}_prof.ncu-rep", "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rd = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rd[0], rd[1], rd[2:]
col = {h: i for i, h in enumerate(hdr)}
with open(f"profiles/{outp}_ncu_full.csv", "w") as f:
    f.write(f"# ncu --set full --clock-control none --import-source on -k regex:{RX} : {CMD}\n")
    f.write("metric,unit," + ",".join('"' + short(r[col["Kernel Name"]]) + '"' for r in data) + "\n")
    for m in METRICS:
        if m in col:
            f.write(f"{m},{units[col[m]]}," + ",".join(r[col[m]].replace(",", "") for r in data) + "\n")
print(open(f"profiles/{outp}_launches.csv").read())
print(open(f"profiles/{outp}_ncu_full.csv").read())
