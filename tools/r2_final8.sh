#!/bin/bash
# final 8-GPU line of round 2 with the default Faddeyeva borders (gpurun --gpus 8 -- bash tools/r2_final8.sh): C2 headline under torchrun
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 $TR --nproc-per-node 8 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 3 --no-cpu > $O/r2_bench_c2_n8_final.json 2> $O/n8.err; echo c2 rc=$?
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2_bench_c2_n8_final.json")); print("c2 n8", d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"], d["farfield_expansion"]["ms_per_step"])
PY
