#!/bin/bash
# 2-GPU line (view-sharded, eager launches, overlapped exchange); bounded so that a hang cannot eat the budget
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 100 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err
python scripts/stage_table.py gpurun_out/bench_2gpu.json | cut -c1-300
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_2gpu.json").read().strip().splitlines()[-1]); print(d["config"].get("launch"), d.get("allreduce"), d["gpu_launches"])
PY
