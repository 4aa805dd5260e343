#!/bin/bash
python -m pytest tests/test_gpu_render.py -x -q --tb=short 2>&1 | tail -4
python bench.py --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err
python bench.py --no-cpu-baseline --no-graph --no-e2e --steps 50 > gpurun_out/bench_eager.json 2>/dev/null
python bench.py --no-cpu-baseline --no-e2e --steps 50 --points 100000 > gpurun_out/bench_100k_graph.json 2>/dev/null
python bench.py --no-cpu-baseline --no-e2e --steps 50 --points 100000 --no-graph > gpurun_out/bench_100k_eager.json 2>/dev/null
python scripts/stage_table.py gpurun_out/bench.json gpurun_out/bench_eager.json gpurun_out/bench_100k_graph.json gpurun_out/bench_100k_eager.json | grep -v "^   " | cut -c1-120; tail -3 gpurun_out/bench.err | cut -c1-300
