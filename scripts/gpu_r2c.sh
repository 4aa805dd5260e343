#!/bin/bash
python -m pytest tests/test_gpu_baseline_parity.py tests/test_gpu_paths.py tests/test_gpu_ops.py tests/test_gpu_render.py -x -q --tb=short 2>&1 | tail -8 > gpurun_out/pytest.log
python scripts/raster_stats.py 2>&1 | head -1 > gpurun_out/raster_stats.txt
python bench.py --steps 50 --no-cpu-baseline --no-e2e > gpurun_out/bench.json 2> gpurun_out/bench.err
for f in 32 256; do DSS_RASTER_FLUSH=$f python bench.py --steps 30 --no-e2e --no-cpu-baseline > gpurun_out/bench_flush$f.json 2>/dev/null; done
tail -8 gpurun_out/pytest.log | cut -c1-300; cat gpurun_out/raster_stats.txt; python scripts/stage_table.py gpurun_out/bench.json gpurun_out/bench_flush32.json gpurun_out/bench_flush256.json | cut -c1-330; tail -2 gpurun_out/bench.err
