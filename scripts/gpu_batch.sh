python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
python scripts/stage_table.py gpurun_out/bench.json; tail -2 gpurun_out/bench.err
