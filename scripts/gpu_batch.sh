python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/pytest.log
python bench.py --steps 50 > gpurun_out/bench.json 2> gpurun_out/bench.err
DSS_OCC_LPS8=1 python bench.py --steps 50 --no-cpu-baseline --no-e2e > gpurun_out/bench_lps8.json 2>> gpurun_out/bench.err
BENCH_E2E_SKIP=h2d python bench.py --steps 50 --no-cpu-baseline > gpurun_out/bench_noh2d.json 2>> gpurun_out/bench.err
BENCH_E2E_SKIP=d2h python bench.py --steps 50 --no-cpu-baseline > gpurun_out/bench_nod2h.json 2>> gpurun_out/bench.err
BENCH_E2E_SKIP=h2d,d2h python bench.py --steps 50 --no-cpu-baseline > gpurun_out/bench_nocopy.json 2>> gpurun_out/bench.err
tail -8 gpurun_out/pytest.log; python scripts/stage_table.py gpurun_out/bench.json gpurun_out/bench_lps8.json gpurun_out/bench_noh2d.json gpurun_out/bench_nod2h.json gpurun_out/bench_nocopy.json
