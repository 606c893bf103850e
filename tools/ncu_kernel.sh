#!/bin/bash
# tools/ncu_kernel.sh <tag> <kernel-regex> [batch] [skip] [count]: one `ncu --set full` capture of the named kernel
set -u
tag=$1; rx=$2; batch=${3:-2000}; skip=${4:-4}; cnt=${5:-2}
out=gpurun_out; mkdir -p $out
cmd="python bench.py --steps 2 --warmup 3 --batch $batch --no-cpu-baseline --no-torch-baseline --no-profile"
timeout 300 $cmd > $out/${tag}_plain.log 2>&1 &&
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"$rx" -s $skip -c $cnt -f -o $out/${tag}_prof $cmd > $out/${tag}_ncu.log 2>&1
echo "ncu rc=$?"; tail -3 $out/${tag}_ncu.log
