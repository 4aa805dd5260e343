#!/bin/bash
# round 2, first GPU check: new rasterizer -- parity tests, work counters, bench
python -m pytest tests -m gpu -x -q --tb=short 2>&1 | tail -15 > gpurun_out/pytest.log
python scripts/raster_stats.py > gpurun_out/raster_stats.txt 2>&1
python bench.py --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -15 gpurun_out/pytest.log | cut -c1-300; cat gpurun_out/raster_stats.txt | tail -8; python scripts/stage_table.py gpurun_out/bench.json; tail -2 gpurun_out/bench.err
