python -m pytest tests -m gpu -x -q --tb=short 2>&1 | tail -5 > gpurun_out/pytest.log
python bench.py --steps 100 --no-cpu-baseline --no-e2e > gpurun_out/bench.json 2> gpurun_out/bench.err
DSS_NO_TILE_ORDER=1 python bench.py --steps 100 --no-cpu-baseline --no-e2e > gpurun_out/bench_noorder.json 2>> gpurun_out/bench.err
tail -3 gpurun_out/pytest.log | cut -c1-200; python scripts/stage_table.py gpurun_out/bench.json gpurun_out/bench_noorder.json; tail -2 gpurun_out/bench.err
