#!/bin/bash
# tools/gpu_new.sh : the statistical-parity / TF32-variant / path-selection tests + a bench line of the TF32 variant
timeout 1000 python -m pytest tests -m gpu -x -q -s -k "statistics or tf32 or path_selection" 2>&1 | grep -v Warn | tail -12
python bench.py --precision tf32 --steps 30 --warmup 3 --no-cpu-baseline --no-torch-baseline 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('tf32 zinc value', round(d['value'],1), d['dtype'], 'fused', round(d['kernel_ms']['fused_layers_kernel']['ms_per_step'],4), 'final', round(d['kernel_ms']['final_edge_kernel']['ms_per_step'],4))
"
