#!/bin/bash
# Runs on the GPU box (through gpurun): parity tests, smoke, a short bench and -- optionally -- the ncu passes.
#   tools/gpu_check.sh <tag> [ncu-regex]
set -u
tag=${1:-run}
rx=${2:-}
out=gpurun_out
mkdir -p $out
(timeout 1200 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_pytest.log)
tail -4 $out/${tag}_pytest.log
(timeout 300 python __graft_entry__.py smoke > $out/${tag}_smoke.log 2>&1; echo "rc=$?" >> $out/${tag}_smoke.log)
tail -2 $out/${tag}_smoke.log
(timeout 600 python bench.py --steps 10 --warmup 3 > $out/${tag}_bench.log 2>&1; echo "rc=$?" >> $out/${tag}_bench.log)
tail -c 600 $out/${tag}_bench.log
if [ -n "$rx" ]; then
  cmd="python bench.py --steps 2 --warmup 3 --batch 2000 --no-cpu-baseline --no-torch-baseline --no-profile"
  timeout 300 $cmd > $out/${tag}_plain.log 2>&1 &&
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/${tag}_launches.csv $cmd > $out/${tag}_ncu1.log 2>&1
  echo "launch list rc=$?"
  timeout 300 $cmd > $out/${tag}_plain2.log 2>&1 &&
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"$rx" -s 20 -c 6 -f -o $out/${tag}_prof $cmd > $out/${tag}_ncu2.log 2>&1
  echo "full set rc=$?"
fi
