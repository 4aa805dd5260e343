python -m pytest tests/test_gpu_ops.py tests/test_gpu_render.py tests/test_gpu_reference_backward.py tests/test_gpu_fullsize.py -m gpu -x -q --tb=short 2>&1 | tail -4 > gpurun_out/pytest.log
python bench.py --steps 60 --no-cpu-baseline --no-e2e > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -3 gpurun_out/pytest.log | cut -c1-200; python scripts/stage_table.py gpurun_out/bench.json; tail -2 gpurun_out/bench.err
