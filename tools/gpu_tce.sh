#!/bin/bash
# tcgen05 per-edge MLP of the per-graph kernel: smoke first (small batch, bounded waits), then parity tests, then the A/B bench
set -u
out=gpurun_out; mkdir -p $out
GDSS_TCEDGE=1 timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/tce_smoke.log 2>&1; rc=$?
tail -3 $out/tce_smoke.log; echo "smoke rc=$rc"
if [ $rc -ne 0 ]; then
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
  exit 0
fi
(GDSS_TCEDGE=1 timeout 1200 python -m pytest tests -m gpu -q -x --timeout 900 > $out/tests.log 2>&1; echo "pytest rc=$?" >> $out/tests.log)
tail -6 $out/tests.log
Q="--no-extra --no-cpu-baseline --no-torch-baseline"
show() { python - "$1" <<'PY'
import json,sys
try:
    txt=open(sys.argv[1]).read(); d=json.loads([l for l in txt.splitlines() if l.startswith('{')][-1])
    km={k:round(v["ms_per_step"],4) for k,v in d.get("kernel_ms",{}).items()}
    print(sys.argv[1], round(d["value"],2), d["unit"], "ms/step", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"],1), km)
except Exception as e: print(sys.argv[1], "failed", e)
PY
}
GDSS_TCEDGE=1 timeout 600 python bench.py $Q > $out/tce_on.json 2> $out/tce_on.err; show $out/tce_on.json
timeout 600 python bench.py $Q > $out/tce_off.json 2> $out/tce_off.err; show $out/tce_off.json
GDSS_TCEDGE=1 timeout 300 python tools/phase_prof.py zinc250k_v2 2960 > $out/phase_zinc_tce.txt 2>&1; cat $out/phase_zinc_tce.txt
