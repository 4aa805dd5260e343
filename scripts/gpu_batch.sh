python -m pytest tests/test_gpu_ops.py -m gpu -x -q --tb=short -k "window_variants or occ_backward_matches" 2>&1 | tail -15 > gpurun_out/pytest.log
tail -15 gpurun_out/pytest.log | cut -c1-220
