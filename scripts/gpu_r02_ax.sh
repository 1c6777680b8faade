#!/bin/bash
# round 2, call AX: confirmation of the final tree (after the removal of the fused Arnoldi sweeps): GPU test tier, smoke, default bench
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout=1200 > gpurun_out/r02ax_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02ax_pytest.log; tail -3 gpurun_out/r02ax_pytest.log; grep -E "^E  |^FAILED|^ERROR" gpurun_out/r02ax_pytest.log | cut -c1-300 | head -20
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
( time python bench.py ) > gpurun_out/r02ax_bench_1gpu.json 2> gpurun_out/r02ax_bench_1gpu.err; echo "bench exit $?"; tail -4 gpurun_out/r02ax_bench_1gpu.err; cut -c1-1200 gpurun_out/r02ax_bench_1gpu.json
