python -m pytest tests/test_gpu_fullsize.py tests/test_knn.py -m gpu -q --tb=short 2>&1 | tail -60 > gpurun_out/pytest_new.log
tail -60 gpurun_out/pytest_new.log | cut -c1-220
