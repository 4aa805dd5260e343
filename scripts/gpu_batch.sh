#!/bin/bash
# Round-end check on a B200 box (run through gpurun from the repo root): parity tests, smoke, bench line, launch list and
# one full ncu capture of each of the two dominant kernels.  Outputs land in gpurun_out/.
python -m pytest tests -m gpu -x -q --tb=short 2>&1 | tail -5 > gpurun_out/pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/b_ncu.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:occ_tile_kernel -c 1 -o gpurun_out/occ_tile -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu1.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:raster_sliced_kernel -c 1 -o gpurun_out/raster -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu2.log 2>&1
tail -3 gpurun_out/pytest.log | cut -c1-200; tail -1 gpurun_out/smoke.log; python scripts/stage_table.py gpurun_out/bench.json; tail -2 gpurun_out/bench.err
