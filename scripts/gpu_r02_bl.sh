#!/bin/bash
# round 2, call BL: final tree - full GPU test tier, smoke, default bench (C3 headline + C5 + C4 + C2)
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout=1200 > gpurun_out/r02bl_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02bl_pytest.log; tail -3 gpurun_out/r02bl_pytest.log; grep -E "^E  |^FAILED|^ERROR" gpurun_out/r02bl_pytest.log | cut -c1-300 | head -20
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
( time python bench.py ) > gpurun_out/r02bl_bench_1gpu.json 2> gpurun_out/r02bl_bench_1gpu.err; echo "bench exit $?"; tail -4 gpurun_out/r02bl_bench_1gpu.err; python -c "
import json; d=json.loads(open('gpurun_out/r02bl_bench_1gpu.json').read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['roofline']['frac']); print(json.dumps(d['secondary_c2'])[:700])"
