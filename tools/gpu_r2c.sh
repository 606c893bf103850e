#!/bin/bash
# round-2 GPU pass: the -m gpu suite, the default bench line, then (each only after its command ran clean without ncu) the ncu
# launch list and one full-set capture of the per-graph kernel and the tcgen05 final-MLP kernel AT THE BENCHED BATCH (B = 10 000)
set -u
out=gpurun_out; mkdir -p $out
(timeout 1500 python -m pytest tests -m gpu -q --timeout 1200 > $out/tests.log 2>&1; echo "pytest rc=$?" >> $out/tests.log)
tail -12 $out/tests.log
(timeout 900 python bench.py > $out/bench.json 2> $out/bench.err; echo "bench rc=$?")
for w in enzymes grid qm9_s4; do timeout 300 python bench.py --workload $w --no-extra --no-cpu-baseline --no-torch-baseline > $out/bench_$w.json 2> $out/bench_$w.err; done
cmd="python bench.py --steps 2 --warmup 3 --burn-in 1 --repeats 1 --no-extra --no-cpu-baseline --no-torch-baseline --no-profile"
timeout 300 $cmd > $out/r02_plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/r02_launches.csv $cmd > $out/r02_ncu1.log 2>&1
echo "launch list rc=$?"
timeout 300 $cmd > $out/r02_plain2.log 2>&1 &&
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"fused_layers_kernel|final_edge_tc_kernel" -s 15 -c 5 -f -o $out/r02_prof $cmd > $out/r02_ncu2.log 2>&1
echo "full set rc=$?"; tail -2 $out/r02_ncu2.log
ls -la $out | head -30
