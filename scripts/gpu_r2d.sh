#!/bin/bash
for im in 0 1; do
  DSS_RASTER_IMM=$im python scripts/raster_stats.py 2>&1 | head -1 > gpurun_out/stats_imm$im.txt
  DSS_RASTER_IMM=$im python bench.py --steps 30 --no-e2e --no-cpu-baseline > gpurun_out/bench_imm$im.json 2>/dev/null
done
DSS_RASTER_IMM=1 DSS_RASTER_FLUSH=32 python bench.py --steps 30 --no-e2e --no-cpu-baseline > gpurun_out/bench_imm1_f32.json 2>/dev/null
DSS_RASTER_IMM=1 DSS_RASTER_FLUSH=32 python scripts/raster_stats.py 2>&1 | head -1 > gpurun_out/stats_imm1_f32.txt
DSS_RASTER_IMM=1 python -m pytest tests/test_gpu_paths.py tests/test_gpu_ops.py -x -q --tb=short 2>&1 | tail -3 > gpurun_out/pytest_imm.log
DSS_RASTER_IMM=1 ncu --set full --import-source on --clock-control none -k regex:raster_sorted -c 1 -o gpurun_out/raster_r2imm -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu2.log 2>&1
for f in imm0 imm1 imm1_f32; do echo $f; cat gpurun_out/stats_$f.txt; python scripts/stage_table.py gpurun_out/bench_$f.json | tail -1 | cut -c1-200; done
cat gpurun_out/pytest_imm.log
