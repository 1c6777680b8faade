#!/bin/bash
# round 2, call AR: split-mode SLQ step (update of one column group under the SpMM of the other): bit identity, then A/B
mkdir -p gpurun_out
python -m pytest tests/test_gpu_spmm.py -m gpu -q --timeout=600 > gpurun_out/r02ar_pytest_spmm.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/r02ar_pytest_spmm.log | cut -c1-300; grep -E "^E  |^FAILED|^ERROR" gpurun_out/r02ar_pytest_spmm.log | cut -c1-300 | head
export KR_BENCH_EDGES=0 KR_BENCH_C4=0
for cfg in "0 1" "1 1" "1 2" "1 3"; do
  set -- $cfg
  echo "== KR_SLQ_SPLIT=$1 KR_SLQ_COMBINE_CTAS=$2"
  KR_SLQ_SPLIT=$1 KR_SLQ_COMBINE_CTAS=$2 python bench.py --steps 4 --warmup 3 2> gpurun_out/r02ar_bench_$1_$2.err | tee gpurun_out/r02ar_bench_$1_$2.json | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
r = d['roofline']
print(json.dumps({'value': round(d['value']), 'ms_per_step': round(d['ms_per_step'], 1), 'e2e': round(d['e2e']['value']), 'frac': round(r['frac'], 4), 'ms_per_launch': round(r['ms_per_launch'], 3), 'cols_per_launch': r.get('columns_per_launch'), 'share': round(r['share_of_step'], 3), 'clocks': d['clocks']['sm_mhz'], 'trace': d['trace_estimate']}))
"
done
