python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/pytest.log
python bench.py --steps 100 > gpurun_out/bench.json 2> gpurun_out/bench.err
DSS_OCC_CONS=10 python bench.py --steps 100 --no-cpu-baseline --no-e2e > gpurun_out/bench_cons10.json 2>> gpurun_out/bench.err
ncu --set full --import-source on --clock-control none -k regex:raster_sliced -c 1 -o gpurun_out/raster -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu2.log 2>&1
tail -3 gpurun_out/pytest.log; python scripts/stage_table.py gpurun_out/bench.json gpurun_out/bench_cons10.json
