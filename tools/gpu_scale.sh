#!/bin/bash
# tools/gpu_scale.sh N : the N-GPU lines (run under gpurun --gpus N): sharded-vs-single parity test, the sampler bench (strong
# scaling = BASELINE config 3, with the weak line as `other_scaling`), the training-step bench
set -u
N=${1:-2}
out=gpurun_out; mkdir -p $out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
if [ "$N" = "2" ]; then
  (timeout 900 python -m pytest tests -m gpu -q -k "multi_gpu" --timeout 900 > $out/tests_multi_gpu.log 2>&1; echo "pytest rc=$?" >> $out/tests_multi_gpu.log)
  tail -3 $out/tests_multi_gpu.log; cat $out/multi_gpu_check.log 2>/dev/null | head -8
fi
(timeout 900 $TR bench.py --gpus $N > $out/bench_${N}gpu.json 2> $out/bench_${N}gpu.err; echo "bench rc=$?")
tail -c 1200 $out/bench_${N}gpu.json
(timeout 600 $TR bench.py --gpus $N --workload zinc_train --steps 5 --warmup 3 --repeats 3 > $out/bench_train_${N}gpu.json 2> $out/bench_train_${N}gpu.err; echo "train bench rc=$?")
tail -c 800 $out/bench_train_${N}gpu.json
