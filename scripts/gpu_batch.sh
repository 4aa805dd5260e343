timeout 420 compute-sanitizer --tool initcheck --error-exitcode 99 --print-limit 15 python -m pytest tests/test_gpu_render.py tests/test_knn.py tests/test_gpu_ops.py -m gpu -q -x -k "test_render_backward_matches_oracle or test_render_forward_matches_oracle or ragged_batches or test_occ_backward_matches_oracle or test_splat_points_matches_oracle" > gpurun_out/initcheck.log 2>&1
echo "initcheck exit $?" >> gpurun_out/initcheck.log
grep -E "ERROR SUMMARY|Uninitialized|exit|passed|failed" gpurun_out/initcheck.log | head -20
grep -A12 "Uninitialized" gpurun_out/initcheck.log | head -60
