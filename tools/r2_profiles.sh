#!/bin/bash
# round-2 profile captures (run on the GPU box: gpurun -- bash tools/r2_profiles.sh); every ncu run follows a plain run of
# the same command that exited 0
set -x
O=gpurun_out
python bench.py --steps 2 --warmup 3 --no-cpu > $O/r2_bench_plain.json 2> $O/r2_bench_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2_launches_bench_c2.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu > $O/ncu_bench.log 2>&1
python tools/k2_profile_target.py 101 direct || exit 1
ncu --set full --import-source on --clock-control none -k regex:line_sum_kernel -c 1 -o $O/r2_k2_direct_full \
    python tools/k2_profile_target.py 101 direct > $O/ncu_k2d.log 2>&1
python tools/k2_profile_target.py 101 expansion || exit 1
ncu --set full --import-source on --clock-control none -k 'regex:line_sum_kernel|farfield_kernel|moments_kernel' -c 3 -o $O/r2_k2_expansion_full \
    python tools/k2_profile_target.py 101 expansion > $O/ncu_k2e.log 2>&1
python tools/k3_profile_target.py 1000000 || exit 1
ncu --set full --import-source on --clock-control none -k regex:table_fit_fused -c 1 -o $O/r2_k3_fused_full \
    python tools/k3_profile_target.py 1000000 > $O/ncu_k3.log 2>&1
python bench.py --workload c5 --steps 20 --warmup 3 --no-cpu --no-graph > $O/r2_bench_c5_plain.json 2> $O/r2_bench_c5_plain.err || exit 1
ncu --set full --import-source on --clock-control none -k regex:rcm_rt_kernel -c 1 -o $O/r2_rcm_rt \
    python bench.py --workload c5 --steps 3 --warmup 3 --no-cpu --no-graph > $O/ncu_c5.log 2>&1
ls -la $O/*.ncu-rep
