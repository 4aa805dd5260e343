#!/bin/bash
# Two-GPU check (gpurun --gpus 2): view-sharded bench with the packed all-reduce.
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 100 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err
python scripts/stage_table.py gpurun_out/bench_2gpu.json; tail -3 gpurun_out/bench_2gpu.err | cut -c1-300
