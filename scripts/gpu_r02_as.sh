#!/bin/bash
# round 2, call AS: full GPU test tier (with the reference-run goldens, the MATLAB drop-in run, the C1 greedy pins and
# the Misc graphs), then the default bench and the reference arm
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout=1200 > gpurun_out/r02as_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02as_pytest.log; tail -3 gpurun_out/r02as_pytest.log; grep -E "^E  |^FAILED|^ERROR" gpurun_out/r02as_pytest.log | cut -c1-300 | head -20
( time python bench.py ) > gpurun_out/r02as_bench_1gpu.json 2> gpurun_out/r02as_bench_1gpu.err; echo "bench exit $?"; tail -4 gpurun_out/r02as_bench_1gpu.err; cut -c1-1200 gpurun_out/r02as_bench_1gpu.json
( time python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/r02as_bench_reference.json 2> gpurun_out/r02as_bench_reference.err; echo "ref exit $?"; tail -4 gpurun_out/r02as_bench_reference.err
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
