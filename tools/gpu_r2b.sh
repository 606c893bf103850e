#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout 1200 > gpurun_out/tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/tests.log
tail -15 gpurun_out/tests.log
python tools/diag_b32.py > gpurun_out/diag_b32.log 2>&1; tail -12 gpurun_out/diag_b32.log
(cd build/r1 && python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-torch-baseline --no-profile) > gpurun_out/bench_r1.json 2> gpurun_out/bench_r1.err
GDSS_NO_PRUNE=1 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-torch-baseline --no-extra > gpurun_out/bench_noprune.json 2> gpurun_out/bench_noprune.err
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-torch-baseline > gpurun_out/bench_new.json 2> gpurun_out/bench_new.err
for f in bench_r1 bench_noprune bench_new; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(f"gpurun_out/{sys.argv[1]}.json").read().strip().splitlines()[-1])
    k=d.get("kernel_ms") or {}
    print(sys.argv[1], "value %.1f ms/step %.3f"%(d["value"], d["ms_per_step"]), {n:round(v["ms_per_step"],3) for n,v in k.items() if v["ms_per_step"]>0.05})
except Exception as e:
    print(sys.argv[1], "failed", e)
PY
done
