#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/replay_weighted_exp_cosh.jsonl
for f in exp cosh; do
timeout 600 python scripts/replay_weighted.py --graphs grid_Mexico --fun $f --hessian --edges 12 --search-space 40 --oracle >> gpurun_out/replay_weighted_exp_cosh.jsonl 2>> gpurun_out/replay_weighted_exp_cosh.err
done
python - <<PY
import json
for l in open('gpurun_out/replay_weighted_exp_cosh.jsonl'):
    d=json.loads(l)
    if 'compare' in d: print(d['graph'], d['method'], 'same_edges', d['same_edges'], 'rel_fval_diff %.1e'%d['rel_fval_diff'], 'speedup run %.2f'%d['speedup_whole_run'])
    else: print('   ', d['fun'], d['impl'], 'time %.2f'%d['time_s'], 'iters', d['iterations'], 'callbacks', d['callbacks'], d['hessian_callbacks'], 'fval', d['fval'])
PY
tail -3 gpurun_out/replay_weighted_exp_cosh.err
