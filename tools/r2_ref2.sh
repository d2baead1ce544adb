#!/bin/bash
# the reference arm under torchrun (OMP_NUM_THREADS=1 exported by the launcher): must finish in minutes and print one line
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
/usr/bin/time -v timeout 600 $TR --nproc-per-node 2 --master-port 29541 bench.py --impl reference --gpus 2 --steps 5 --warmup 3 > $O/r2_bench_reference_n2.json 2> $O/ref2.err
echo rc=$?; grep -E "Elapsed|Maximum resident" $O/ref2.err; cat $O/r2_bench_reference_n2.json | cut -c1-600
timeout 300 $TR --nproc-per-node 2 --master-port 29542 bench.py --gpus 2 --steps 5 --warmup 3 > $O/r2_bench_c2_n2_b.json 2> $O/n2b.err; echo rc=$?
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_c2_n2_b.json')); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['frac'], d['farfield_expansion']['ms_per_step'])"
