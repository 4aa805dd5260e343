python -m pytest tests -m gpu -x -q --tb=short 2>&1 | tail -8 > gpurun_out/pytest.log
python bench.py --steps 100 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err
DSS_OCC_WIDE=1 python -m pytest tests/test_gpu_render.py tests/test_gpu_ops.py tests/test_gpu_reference_backward.py -m gpu -x -q --tb=short 2>&1 | tail -4 >> gpurun_out/pytest.log
DSS_OCC_WIDE=1 python bench.py --steps 100 --no-cpu-baseline --no-e2e > gpurun_out/bench_wide.json 2>> gpurun_out/bench.err
tail -10 gpurun_out/pytest.log | cut -c1-200; python scripts/stage_table.py gpurun_out/bench.json gpurun_out/bench_wide.json; tail -2 gpurun_out/bench.err
