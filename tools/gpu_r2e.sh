#!/bin/bash
# round-2 pass: full -m gpu suite, default bench, small-shard A/B of the bucket stream priorities, C4 lines
set -u
out=gpurun_out; mkdir -p $out
(timeout 1500 python -m pytest tests -m gpu -q --timeout 1200 > $out/tests.log 2>&1; echo "pytest rc=$?" >> $out/tests.log)
tail -4 $out/tests.log
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
for b in 1250 2500; do
  timeout 300 python bench.py --batch $b $Q > $out/shard_${b}_prio.json 2> $out/shard_${b}_prio.err; show $out/shard_${b}_prio.json
  GDSS_NO_PRIO=1 timeout 300 python bench.py --batch $b $Q > $out/shard_${b}_noprio.json 2> $out/shard_${b}_noprio.err; show $out/shard_${b}_noprio.json
done
for w in enzymes grid; do timeout 300 python bench.py --workload $w $Q > $out/bench_$w.json 2> $out/bench_$w.err; show $out/bench_$w.json; done
timeout 900 python bench.py > $out/bench.json 2> $out/bench.err; echo "bench rc=$?"; show $out/bench.json
