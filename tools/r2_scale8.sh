#!/bin/bash
# 8-GPU evidence for round 2 (gpurun --gpus 8 -- bash tools/r2_scale8.sh): C2 headline, C5 with the fused exchange and with NCCL
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 $TR --nproc-per-node 8 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 3 > $O/r2_bench_c2_n8.json 2> $O/n8.err; echo c2 rc=$?
timeout 300 $TR --nproc-per-node 8 --master-port 29523 bench.py --gpus 8 --workload c5 --steps 200 --warmup 5 --no-cpu > $O/r2_bench_c5_n8_peer.json 2> $O/c5n8p.err; echo c5 peer rc=$?
timeout 300 $TR --nproc-per-node 8 --master-port 29524 bench.py --gpus 8 --workload c5 --steps 200 --warmup 5 --no-cpu --c5-collective nccl > $O/r2_bench_c5_n8_nccl.json 2> $O/c5n8n.err; echo c5 nccl rc=$?
python - <<'PY'
import json
for f in ("r2_bench_c2_n8", "r2_bench_c5_n8_peer", "r2_bench_c5_n8_nccl"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
        print(f, d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"], d["roofline"]["frac"], d.get("farfield_expansion", {}).get("ms_per_step"))
    except Exception as e:
        print(f, "no line", e)
PY
