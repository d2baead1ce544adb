#!/bin/bash
# the reference arm under torchrun (the launcher exports OMP_NUM_THREADS=1 for nproc > 1): must finish in minutes and print one
# line with all host cores in use.  No CUDA call on this arm, so a one-GPU box serves.
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
t0=$(date +%s)
timeout 900 $TR --nproc-per-node 2 --master-port 29541 bench.py --impl reference --gpus 2 --steps 5 --warmup 3 > $O/r2_bench_reference_n2.json 2> $O/ref2.err
echo rc=$? seconds=$(( $(date +%s) - t0 ))
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_reference_n2.json')); print(d['impl'], d['value'], d['ms_per_step'], d['cpu_baseline']['cores'], d['config']['workload'], d['n_gpus'])"
