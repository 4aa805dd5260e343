python -m pytest tests -m gpu -x -q --tb=short 2>&1 | tail -5 > gpurun_out/pytest.log
python bench.py --steps 100 > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -3 gpurun_out/pytest.log | cut -c1-200; python scripts/stage_table.py gpurun_out/bench.json; tail -2 gpurun_out/bench.err
