#!/bin/bash
# full ncu captures of the three binning kernels (one launch each)
for k in bin_count_kernel bin_scatter_kernel occ_cellbin_kernel; do
  ncu --set full --import-source on --clock-control none -k regex:$k -c 1 -o gpurun_out/r2_$k -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_$k.log 2>&1
done
ls -la gpurun_out/*.ncu-rep | tail -5
