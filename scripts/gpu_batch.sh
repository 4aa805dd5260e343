python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/pytest.log
python bench.py --steps 100 > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -3 gpurun_out/pytest.log; python scripts/stage_table.py gpurun_out/bench.json
