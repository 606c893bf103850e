#!/bin/bash
# A/B on one box: previous build (build/ab/libgdss_old.so) vs the current one with / without the tcgen05 per-edge MLP
set -u
out=gpurun_out; mkdir -p $out
Q="--no-extra --no-cpu-baseline --no-torch-baseline ${BENCH_ARGS:-}"
show() { python - "$1" <<'PY'
import json,sys
try:
    txt=open(sys.argv[1]).read(); d=json.loads([l for l in txt.splitlines() if l.startswith('{')][-1])
    km={k:round(v["ms_per_step"],4) for k,v in d.get("kernel_ms",{}).items()}
    print(sys.argv[1], round(d["value"],2), d["unit"], "ms/step", round(d["ms_per_step"],4), "fused", km.get("fused_layers_kernel"))
except Exception as e: print(sys.argv[1], "failed", e)
PY
}
GDSS_LIB=$PWD/build/ab/libgdss_old.so timeout 600 python bench.py $Q > $out/ab_old.json 2> $out/ab_old.err; show $out/ab_old.json
timeout 600 python bench.py $Q > $out/ab_off.json 2> $out/ab_off.err; show $out/ab_off.json
GDSS_TCEDGE=1 timeout 600 python bench.py $Q > $out/ab_on.json 2> $out/ab_on.err; show $out/ab_on.json
for s in 2; do GDSS_TCEDGE=1 GDSS_TCE_SLOTS=$s timeout 600 python bench.py $Q > $out/ab_on_s$s.json 2> $out/ab_on_s$s.err; show $out/ab_on_s$s.json; done
