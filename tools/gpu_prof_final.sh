#!/bin/bash
# final ncu pass of the round (each capture only after its command ran clean without ncu): sampler at the benched batch, then the
# training step
set -u
out=gpurun_out; mkdir -p $out
cmd="python bench.py --steps 2 --warmup 3 --burn-in 1 --repeats 1 --no-extra --no-cpu-baseline --no-torch-baseline --no-profile"
timeout 300 $cmd > $out/r02f_plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/r02f_launches.csv $cmd > $out/r02f_ncu1.log 2>&1
echo "sampler launch list rc=$?"
timeout 300 $cmd > $out/r02f_plain2.log 2>&1 &&
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"fused_layers_kernel|final_edge_tc_kernel" -s 15 -c 5 -f -o $out/r02f_prof $cmd > $out/r02f_ncu2.log 2>&1
echo "sampler full set rc=$?"; tail -2 $out/r02f_ncu2.log
tcmd="python bench.py --workload zinc_train --steps 1 --warmup 1 --repeats 1 --no-torch-baseline --no-cpu-baseline"
timeout 300 $tcmd > $out/r02t_plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $out/r02t_launches.csv $tcmd > $out/r02t_ncu1.log 2>&1
echo "train launch list rc=$?"
timeout 300 $tcmd > $out/r02t_plain2.log 2>&1 &&
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"tr_linear_kernel|tr_wgrad_kernel|tr_gcn_bwd_kernel|tr_gcn_fwd_kernel|tr_maps_bwd_kernel" -s 300 -c 10 -f -o $out/r02t_prof $tcmd > $out/r02t_ncu2.log 2>&1
echo "train full set rc=$?"; tail -2 $out/r02t_ncu2.log
ls -la $out | grep r02
