#!/bin/bash
# round-end check without the ncu captures: parity tests, smoke, the bench line (our arm) and the reference arm
python -m pytest tests -m gpu -x -q --tb=short 2>&1 | tail -5 > gpurun_out/pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -3 gpurun_out/pytest.log | cut -c1-200; tail -1 gpurun_out/smoke.log; python scripts/stage_table.py gpurun_out/bench.json; tail -2 gpurun_out/bench.err
