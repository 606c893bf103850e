#!/bin/bash
# tools/ncu_qm9.sh <tag>: bench line + one `ncu --set full` capture of the per-graph kernel on C2 (QM9, S4, B = 10000)
set -u
tag=$1; out=gpurun_out; mkdir -p $out
cmd="python bench.py --workload qm9_s4 --steps 2 --warmup 3 --no-cpu-baseline --no-torch-baseline --no-profile"
bash tools/gpu_ab.sh
timeout 300 $cmd > $out/${tag}_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fused_layers_kernel -s 6 -c 1 -f -o $out/${tag}_prof $cmd > $out/${tag}_ncu.log 2>&1
echo "ncu rc=$?"; tail -2 $out/${tag}_ncu.log
