#!/bin/bash
python -m pytest tests/test_gpu_ops.py tests/test_gpu_paths.py tests/test_gpu_baseline_parity.py tests/test_gpu_fullsize.py -x -q --tb=short 2>&1 | tail -4 > gpurun_out/pytest.log
python bench.py --steps 50 --no-e2e --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -3 gpurun_out/pytest.log | cut -c1-200; python scripts/stage_table.py gpurun_out/bench.json | cut -c1-300; tail -2 gpurun_out/bench.err
