#!/bin/bash
# Round-end evidence on ONE GPU: test suite, bench line, launch list, ncu captures of the committed kernels.
#   gpurun --timeout 1500 -- 'bash tools/final_profile.sh'
# Everything lands in gpurun_out/; tools/ncu_summary.py + tools/launch_summary.py condense it into profiles/.
set -u
o=gpurun_out; mkdir -p $o
python -m pytest tests -m gpu -q > $o/final_gpu_tests.log 2>&1; tail -2 $o/final_gpu_tests.log
python bench.py > $o/r02_bench_c3_n1.json 2> $o/bench_err.log; cut -c1-260 $o/r02_bench_c3_n1.json
python bench.py --workload c2 --no-cpu-baseline > $o/r02_bench_c2_n1.json 2>> $o/bench_err.log; cut -c1-200 $o/r02_bench_c2_n1.json
python tools/dense_probe.py --json $o/r02_bench_c2_dense.json > /dev/null 2>> $o/bench_err.log
B="python bench.py --steps 2 --warmup 1 --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $o/r02_launches_c3.csv $B > $o/ncu_launches.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:los_rl_kernel -s 2 -c 1 -o $o/r02_los_final -f $B > $o/ncu_los.log 2>&1
ncu --set full --clock-control none -k regex:field_b_il_kernel -s 2 -c 1 -o $o/r02_field_final -f $B > $o/ncu_field.log 2>&1
ncu --set full --clock-control none -k regex:bg_gram64_kernel -s 4 -c 2 -o $o/r02_gram_final -f $B > $o/ncu_gram.log 2>&1
D="python tools/dense_probe.py --once"
ncu --set full --clock-control none -k regex:build_cov_kernel -s 1 -c 1 -o $o/r02_dense_build_cov -f $D > $o/ncu_dense1.log 2>&1
ncu --set full --clock-control none -k regex:chol_update128_kernel -s 9 -c 1 -o $o/r02_dense_update128 -f $D > $o/ncu_dense2.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $o/r02_launches_dense.csv $D > /dev/null 2>&1
ls -la $o/*.ncu-rep | tail -8
