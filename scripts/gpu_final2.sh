#!/bin/bash
bash scripts/gpu_final.sh
python bench.py --steps 40 --no-cpu-baseline --colours view > gpurun_out/bench_colours_view.json 2>/dev/null
python bench.py --steps 40 --no-cpu-baseline --colours shaded > gpurun_out/bench_colours_shaded.json 2>/dev/null
python scripts/stage_table.py gpurun_out/bench_colours_view.json gpurun_out/bench_colours_shaded.json | cut -c1-330
