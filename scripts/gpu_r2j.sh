#!/bin/bash
python -m pytest tests -m gpu -x -q --tb=short 2>&1 | tail -6 > gpurun_out/pytest.log
python scripts/raster_stats.py 2>&1 | head -1 > gpurun_out/raster_stats.txt
python bench.py --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err
DSS_RASTER_MINB=5 python bench.py --steps 40 --no-e2e --no-cpu-baseline > gpurun_out/bench_minb5.json 2>/dev/null
tail -6 gpurun_out/pytest.log | cut -c1-300; cat gpurun_out/raster_stats.txt; python scripts/stage_table.py gpurun_out/bench.json gpurun_out/bench_minb5.json | cut -c1-330; tail -2 gpurun_out/bench.err
