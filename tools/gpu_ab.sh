#!/bin/bash
# tools/gpu_ab.sh : parity tests + short ZINC / QM9 bench lines (value, fused-kernel ms) on the GPU box
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2)
for w in qm9_s4 zinc250k_v2; do
  python bench.py --workload $w --steps 30 --warmup 3 --no-cpu-baseline --no-torch-baseline 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('$w value', round(d['value'],1), 'fused', round(d['kernel_ms']['fused_layers_kernel']['ms_per_step'],4), 'final', round(d['kernel_ms']['final_edge_kernel']['ms_per_step'],4))
"
done
