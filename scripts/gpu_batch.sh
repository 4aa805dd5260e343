python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/pytest.log
python bench.py --steps 50 > gpurun_out/bench.json 2> gpurun_out/bench.err
DSS_BIN_DIRECT=1 python bench.py --steps 50 --no-cpu-baseline --no-e2e > gpurun_out/bench_bindirect.json 2>> gpurun_out/bench.err
python scripts/raster_stats.py > gpurun_out/raster_stats.txt 2>&1
ncu --set full --import-source on --clock-control none -k regex:occ_tile_kernel -c 1 -o gpurun_out/occ_tile -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu1.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:raster_sliced -c 1 -o gpurun_out/raster -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu2.log 2>&1
tail -8 gpurun_out/pytest.log; cat gpurun_out/bench.json; cat gpurun_out/bench_bindirect.json
