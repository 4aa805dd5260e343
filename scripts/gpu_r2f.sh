#!/bin/bash
python -m pytest tests/test_gpu_render.py -x -q --tb=short 2>&1 | tail -5 > gpurun_out/pytest.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 100 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err
tail -5 gpurun_out/pytest.log | cut -c1-300
python scripts/stage_table.py gpurun_out/bench_2gpu.json | cut -c1-400; tail -3 gpurun_out/bench_2gpu.err | cut -c1-300
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_2gpu.json').read().strip().splitlines()[-1])
print("allreduce", d.get("allreduce")); print("e2e", d.get("e2e"))
PY
