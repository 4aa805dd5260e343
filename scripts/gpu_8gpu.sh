#!/bin/bash
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 100 --warmup 3 > gpurun_out/bench_8gpu.json 2> gpurun_out/bench_8gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 8 --steps 40 --warmup 3 --image-size 1024 > gpurun_out/bench_c4_8gpu_1024.json 2> gpurun_out/bench_c4.err
python scripts/stage_table.py gpurun_out/bench_8gpu.json gpurun_out/bench_c4_8gpu_1024.json | cut -c1-400
python - <<'PY'
import json
for f in ("gpurun_out/bench_8gpu.json", "gpurun_out/bench_c4_8gpu_1024.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1]); print(f, "allreduce", d.get("allreduce"), "e2e", (d.get("e2e") or {}).get("value"))
    except Exception as e: print(f, e)
PY
tail -2 gpurun_out/bench_8gpu.err | cut -c1-300; tail -2 gpurun_out/bench_c4.err | cut -c1-300
