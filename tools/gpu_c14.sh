#!/bin/bash
# tools/gpu_c14.sh : short bench lines of the non-headline BASELINE configs (C1 community_small, C4 ENZYMES / grid)
for w in enzymes grid community_small; do
  timeout 280 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline --no-torch-baseline > gpurun_out/c4_$w.log 2>&1
  python - $w <<'P'
import sys, json
for l in open(f'gpurun_out/c4_{sys.argv[1]}.log'):
    if l.startswith('{'):
        d = json.loads(l)
        print(d['config']['workload'], 'B', d['config']['batch'], 'value', round(d['value'], 2), 'ms/step', round(d['ms_per_step'], 3))
        print('   ', ' '.join(f"{k}={round(v['ms_per_step'], 4)}" for k, v in d['kernel_ms'].items()))
P
done
