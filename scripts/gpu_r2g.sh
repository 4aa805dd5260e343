#!/bin/bash
python -m pytest tests/test_shading.py tests/test_losses.py tests/test_gpu_render.py tests/test_gpu_api.py -m gpu -x -q --tb=short 2>&1 | tail -12 > gpurun_out/pytest.log
python bench.py --steps 50 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -12 gpurun_out/pytest.log | cut -c1-300; python scripts/stage_table.py gpurun_out/bench.json | cut -c1-400; tail -2 gpurun_out/bench.err
