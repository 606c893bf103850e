#!/bin/bash
# final check of the round on one GPU: smoke, the whole -m gpu suite, the default bench line (each bounded by its own timeout)
set -u
out=gpurun_out; mkdir -p $out
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $out/smoke.log
(timeout 400 python -m pytest tests -m gpu -q --timeout 300 > $out/tests.log 2>&1; echo "pytest rc=$?" >> $out/tests.log); tail -3 $out/tests.log
timeout 300 python bench.py > $out/bench.json 2> $out/bench.err; echo "bench rc=$?"
python - <<'PY'
import json
txt=open("gpurun_out/bench.json").read(); d=json.loads([l for l in txt.splitlines() if l.startswith('{')][-1])
print(round(d["value"],1), d["unit"], round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), "eager x", d["torch_eager_same_gpu"]["speedup_of_value"], {k: round(v["value"],2) for k,v in d["extra"].items()})
PY
