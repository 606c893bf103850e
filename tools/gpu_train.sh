#!/bin/bash
# training-step pass: the DSM parity tests, the training bench line, its ncu launch list
set -u
out=gpurun_out; mkdir -p $out
(timeout 900 python -m pytest tests -m gpu -q -k "dsm or training or optimizer" --timeout 600 > $out/tests_train.log 2>&1; echo "pytest rc=$?" >> $out/tests_train.log)
tail -5 $out/tests_train.log
timeout 600 python bench.py --workload zinc_train --steps 5 --warmup 3 --repeats 3 > $out/bench_train.json 2> $out/bench_train.err; echo "train bench rc=$?"
python - <<PY
import json
d=json.load(open("$out/bench_train.json")); print(d["value"], d["unit"], d["ms_per_step"], "eager", d.get("torch_eager_same_gpu"))
PY
if [ "${1:-}" = "ncu" ]; then
cmd="python bench.py --workload zinc_train --steps 1 --warmup 1 --repeats 1 --no-torch-baseline"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $out/r02_train_launches.csv $cmd > $out/train_ncu.log 2>&1
echo "train launch list rc=$?"
fi
