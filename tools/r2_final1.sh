#!/bin/bash
# final 1-GPU evidence of round 2 (gpurun -- bash tools/r2_final1.sh): smoke, default bench line, launch list, full ncu captures of
# the two launches of the direct Voigt line sum (every ncu run follows a plain run of the same command that exited 0)
O=gpurun_out
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py > $O/r2_bench_final_n1.json 2> $O/r2_bench_final_n1.err || { tail -5 $O/r2_bench_final_n1.err; exit 1; }
python bench.py --steps 2 --warmup 3 --no-cpu > $O/r2_bench_plain.json 2> $O/r2_bench_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2_launches_bench_c2_final.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu > $O/ncu_bench.log 2>&1
python tools/k2_profile_target.py 101 direct || exit 1
ncu --set full --import-source on --clock-control none -k "regex:line_sum_kernel|far_fold_kernel" -c 2 -o $O/r2_k2_split_final \
    python tools/k2_profile_target.py 101 direct > $O/ncu_k2split.log 2>&1
python bench.py --workload c5 --steps 200 --warmup 5 > $O/r2_bench_c5_final_n1.json 2> $O/r2_bench_c5_final_n1.err
python -c "
import json
d=json.load(open('$O/r2_bench_final_n1.json')); print('c2', d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_ms_per_step'], d['farfield_expansion']['ms_per_step'], d['parity_sample']['ok'])
d=json.load(open('$O/r2_bench_c5_final_n1.json')); print('c5', d['ms_per_step'], d['value'], d['roofline']['frac'], d.get('parity_sample',{}).get('ok'))"
