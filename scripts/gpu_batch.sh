python -m pytest tests -m gpu -x -q --tb=short 2>&1 | tail -5 > gpurun_out/pytest.log
python bench.py --steps 100 > gpurun_out/bench.json 2> gpurun_out/bench.err
DSS_RASTER_MINB=5 python bench.py --steps 100 --no-cpu-baseline --no-e2e > gpurun_out/bench_minb5.json 2>> gpurun_out/bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/b_ncu.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:occ_tile_kernel -c 1 -o gpurun_out/occ_tile -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu1.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:raster_sliced -c 1 -o gpurun_out/raster -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu2.log 2>&1
tail -3 gpurun_out/pytest.log | cut -c1-200; python scripts/stage_table.py gpurun_out/bench.json gpurun_out/bench_minb5.json; tail -2 gpurun_out/bench.err
