#!/bin/bash
# 4-GPU evidence for round 2 (gpurun --gpus 4 -- bash tools/r2_scale4.sh)
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
python -m pytest tests -x -q -m gpu -k "distinct or device_group or sharded_tables" 2>&1 | tail -3 > $O/r2_group_tests_4gpu.log; cat $O/r2_group_tests_4gpu.log
$TR --nproc-per-node 4 --master-port 29522 bench.py --gpus 4 --steps 20 --warmup 3 > $O/r2_bench_c2_n4.json 2> $O/n4.err; tail -2 $O/n4.err
$TR --nproc-per-node 4 --master-port 29523 bench.py --gpus 4 --workload c5 --steps 200 --warmup 5 > $O/r2_bench_c5_n4.json 2> $O/c5n4.err; tail -2 $O/c5n4.err
python bench.py --single-process --gpus 4 --steps 10 > $O/r2_bench_c2_single_process_n4.json 2> $O/sp4.err; tail -2 $O/sp4.err
