#!/bin/bash
for v in "A:" "B:DSS_BIN_NOTAB=1" "C:DSS_BIN_NORECTS=1" "D:DSS_BIN_NOTAB=1 DSS_BIN_NORECTS=1"; do
  name=${v%%:*}; envs=${v#*:}
  env $envs python bench.py --steps 40 --no-e2e --no-cpu-baseline > gpurun_out/bench_bin$name.json 2>/dev/null
  echo "$name [$envs]"; python scripts/stage_table.py gpurun_out/bench_bin$name.json | tail -1 | cut -c1-230
done
