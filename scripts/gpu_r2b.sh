#!/bin/bash
# sweep of the rasterizer's flush threshold + one full ncu capture of the raster kernel
for f in 96 256 512 768; do
  DSS_RASTER_FLUSH=$f python bench.py --steps 30 --no-e2e --no-cpu-baseline > gpurun_out/bench_flush$f.json 2> gpurun_out/bench_flush$f.err
  DSS_RASTER_FLUSH=$f python scripts/raster_stats.py 2>&1 | head -1 > gpurun_out/stats_flush$f.txt
done
ncu --set full --import-source on --clock-control none -k regex:raster_sorted -c 1 -o gpurun_out/raster_r2 -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu2.log 2>&1
python -m pytest tests/test_gpu_baseline_parity.py tests/test_gpu_paths.py tests/test_gpu_reference_backward.py -x -q --tb=short 2>&1 | tail -8 > gpurun_out/pytest.log
for f in 96 256 512 768; do echo "flush $f"; cat gpurun_out/stats_flush$f.txt; python scripts/stage_table.py gpurun_out/bench_flush$f.json | cut -c1-400; done
tail -8 gpurun_out/pytest.log | cut -c1-300
