#!/bin/bash
# round-2 experiments: tiny-graph occupancy variants on C2, C4 lines after the final_node change, launch list of the training step
set -u
out=gpurun_out; mkdir -p $out
Q="--no-extra --no-cpu-baseline --no-torch-baseline"
for occ in 4 5 6 8; do
  GDSS_TINY_OCC=$occ timeout 300 python bench.py --workload qm9_s4 $Q > $out/qm9_occ$occ.json 2> $out/qm9_occ$occ.err
  python - <<PY
import json
try:
    d=json.load(open("$out/qm9_occ$occ.json")); print("qm9 occ $occ:", round(d["value"]), "graphs/s", d["ms_per_step"], "fused", d["kernel_ms"]["fused_layers_kernel"]["ms_per_step"])
except Exception as e: print("qm9 occ $occ failed", e)
PY
done
for w in enzymes grid; do timeout 300 python bench.py --workload $w $Q > $out/bench_$w.json 2> $out/bench_$w.err
  python - <<PY
import json
try:
    d=json.load(open("$out/bench_$w.json")); print("$w:", d["value"], "graphs/s", d["ms_per_step"], {k:round(v["ms_per_step"],4) for k,v in d["kernel_ms"].items()})
except Exception as e: print("$w failed", e)
PY
done
cmd="python bench.py --workload zinc_train --steps 1 --warmup 1 --repeats 1 --no-torch-baseline"
timeout 300 $cmd > $out/train_plain.log 2>&1; echo "train plain rc=$?"; tail -c 300 $out/train_plain.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $out/r02_train_launches.csv $cmd > $out/train_ncu.log 2>&1
echo "train launch list rc=$?"
