#!/bin/bash
# 8-GPU headline line (bounded); add "c4" as first argument for BASELINE config 4 (1024^2) as well
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 100 --warmup 3 > gpurun_out/bench_8gpu.json 2> gpurun_out/bench_8gpu.err
if [ "$1" = "c4" ]; then
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 8 --steps 40 --warmup 3 --image-size 1024 > gpurun_out/bench_c4_8gpu_1024.json 2> gpurun_out/bench_c4.err
fi
python scripts/stage_table.py gpurun_out/bench_8gpu.json | cut -c1-300
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_8gpu.json").read().strip().splitlines()[-1]); print("allreduce", d.get("allreduce"), "e2e", (d.get("e2e") or {}).get("value"))
PY
