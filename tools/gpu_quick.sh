#!/bin/bash
# tools/gpu_quick.sh <tag>: parity tests + a short bench (no CPU / torch baselines) on the GPU box
set -u
tag=${1:-q}
out=gpurun_out
mkdir -p $out
(timeout 900 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_pytest.log)
tail -15 $out/${tag}_pytest.log
(timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-torch-baseline > $out/${tag}_bench.log 2>&1; echo "rc=$?" >> $out/${tag}_bench.log)
python - <<PY
import json
for l in open("$out/${tag}_bench.log"):
    if l.startswith("{"):
        d = json.loads(l)
        print("value", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"], "clk", d["clocks"]["sm_mhz"])
        for k, v in (d.get("kernel_ms") or {}).items():
            print(f"   {k:22s} {v['ms_per_step']:.3f} ms/step {v['share']*100:.1f}%")
PY
tail -2 $out/${tag}_bench.log | cut -c1-300
