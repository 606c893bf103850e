#!/bin/bash
# round-2 GPU check: the whole -m gpu suite, smoke, the default bench line (all under gpurun_out/)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt 2>&1
python -m pytest tests -m gpu -q -x --timeout 1200 > gpurun_out/tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/tests.log
tail -5 gpurun_out/tests.log
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log; tail -2 gpurun_out/smoke.log
python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/bench.json
