#!/bin/bash
# quick check of a per-graph-kernel change: score / sampler parity tests, then the default bench line without the extras
set -u
out=gpurun_out; mkdir -p $out
(timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 -k "score_networks or sampler_steps or alternate or degenerate or full_1000 or op_level" > $out/tests_quick.log 2>&1; echo "pytest rc=$?" >> $out/tests_quick.log)
tail -4 $out/tests_quick.log
timeout 600 python bench.py --no-extra --no-cpu-baseline --no-torch-baseline ${BENCH_ARGS:-} > $out/quick.json 2> $out/quick.err
python - <<'PY'
import json
txt=open("gpurun_out/quick.json").read(); d=json.loads([l for l in txt.splitlines() if l.startswith('{')][-1])
print(round(d["value"],2), d["unit"], "ms/step", round(d["ms_per_step"],4), {k:round(v["ms_per_step"],4) for k,v in d["kernel_ms"].items()})
PY
