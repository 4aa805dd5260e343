#!/bin/bash
python -m pytest tests/test_gpu_ops.py -q -x --tb=short 2>&1 | tail -3
python bench.py --steps 40 --no-e2e --no-cpu-baseline > gpurun_out/bench_base.json 2>/dev/null
echo base; python scripts/stage_table.py gpurun_out/bench_base.json | tail -1 | cut -c1-230
for f in dss_b200/lib/variants/*.so; do
  n=$(basename $f .so)
  DSS_B200_LIB=$PWD/$f python bench.py --steps 40 --no-e2e --no-cpu-baseline > gpurun_out/bench_$n.json 2>/dev/null
  echo $n; python scripts/stage_table.py gpurun_out/bench_$n.json | tail -1 | cut -c1-130
done
