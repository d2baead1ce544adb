#!/bin/bash
# 4-GPU evidence for round 2 (gpurun --gpus 4 -- bash tools/r2_scale4.sh): C2 at N = 4 and N = 2, C5 at N = 4 (fused exchange)
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 $TR --nproc-per-node 4 --master-port 29522 bench.py --gpus 4 --steps 20 --warmup 3 --no-cpu > $O/r2_bench_c2_n4.json 2> $O/n4.err; echo c2n4 rc=$?
timeout 300 $TR --nproc-per-node 2 --master-port 29525 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu > $O/r2_bench_c2_n2.json 2> $O/n2.err; echo c2n2 rc=$?
timeout 300 $TR --nproc-per-node 4 --master-port 29523 bench.py --gpus 4 --workload c5 --steps 200 --warmup 5 --no-cpu > $O/r2_bench_c5_n4.json 2> $O/c5n4.err; echo c5n4 rc=$?
timeout 300 python -m pytest tests -x -q -m gpu -k "distinct or device_group or sharded_tables or fused_peer" 2>&1 | tail -3 > $O/r2_group_tests_4gpu.log; cat $O/r2_group_tests_4gpu.log
python - <<'PY'
import json
for f in ("r2_bench_c2_n4", "r2_bench_c2_n2", "r2_bench_c5_n4"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
        print(f, d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"], d["roofline"]["frac"], d.get("farfield_expansion", {}).get("ms_per_step"))
    except Exception as e:
        print(f, "no line", e)
PY
