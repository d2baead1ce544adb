#!/bin/bash
# 2-GPU check of the fused RCM step (gpurun --gpus 2 -- bash tools/r2_peer2.sh): group tests, then C5 under torchrun with the
# peer-memory exchange and with the NCCL all-reduce
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 python -m pytest tests -x -q -m gpu -k "distinct or device_group or sharded_tables" 2>&1 | tail -5 > $O/r2_group_tests_2gpu_peer.log; cat $O/r2_group_tests_2gpu_peer.log
timeout 300 $TR --nproc-per-node 2 --master-port 29531 bench.py --gpus 2 --workload c5 --steps 200 --warmup 5 > $O/r2_bench_c5_n2_peer.json 2> $O/c5n2p.err; tail -3 $O/c5n2p.err
timeout 300 $TR --nproc-per-node 2 --master-port 29532 bench.py --gpus 2 --workload c5 --steps 200 --warmup 5 --c5-collective nccl --no-cpu > $O/r2_bench_c5_n2_nccl.json 2> $O/c5n2n.err; tail -3 $O/c5n2n.err
python - <<'PY'
import json
for f in ("r2_bench_c5_n2_peer", "r2_bench_c5_n2_nccl"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
        print(f, d["ms_per_step"], d["value"], d.get("collective"), d.get("parity_sample", {}).get("ok"))
    except Exception as e:
        print(f, "no line", e)
PY
