#!/bin/bash
# A/B of the TF32 split variants: error probe + short bench lines with the default library, then with libgdss_b200_alt.so
python tools/err_probe.py 2>&1 | grep -v Warn
for w in zinc250k_v2 qm9_s4; do python bench.py --workload $w --steps 30 --warmup 3 --no-cpu-baseline --no-torch-baseline 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('$w value', round(d['value'],1), 'fused', round(d['kernel_ms']['fused_layers_kernel']['ms_per_step'],4), 'final', round(d['kernel_ms']['final_edge_kernel']['ms_per_step'],4))
"; done
echo ---- alt
cp gdss_b200/csrc/libgdss_b200_alt.so gdss_b200/csrc/libgdss_b200.so
python tools/err_probe.py 2>&1 | grep -v Warn
for w in zinc250k_v2 qm9_s4; do python bench.py --workload $w --steps 30 --warmup 3 --no-cpu-baseline --no-torch-baseline 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('$w value', round(d['value'],1), 'fused', round(d['kernel_ms']['fused_layers_kernel']['ms_per_step'],4), 'final', round(d['kernel_ms']['final_edge_kernel']['ms_per_step'],4))
"; done
