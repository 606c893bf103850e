#!/bin/bash
set -u
out=gpurun_out; mkdir -p $out
(timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 -k "enzymes or grid or general or degenerate or score_networks" > $out/tests_c4.log 2>&1; echo "pytest rc=$?" >> $out/tests_c4.log)
tail -4 $out/tests_c4.log
for w in enzymes grid; do timeout 300 python bench.py --workload $w --no-extra --no-cpu-baseline --no-torch-baseline > $out/bench_$w.json 2> $out/bench_$w.err
python - $w <<'PY'
import json,sys
w=sys.argv[1]
txt=open(f"gpurun_out/bench_{w}.json").read(); d=json.loads([l for l in txt.splitlines() if l.startswith('{')][-1])
print(w, round(d["value"],2), d["unit"], "ms/step", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"],2))
PY
done
