#!/bin/bash
mkdir -p gpurun_out
timeout 900 python scripts/replay_weighted.py --graphs grid_England,grid_Mexico --hessian --edges 12 --search-space 40 --oracle > gpurun_out/replay_weighted_hessian.jsonl 2> gpurun_out/replay_weighted_hessian.err
python - <<PY
import json
for l in open('gpurun_out/replay_weighted_hessian.jsonl'):
    d=json.loads(l)
    if 'compare' in d: print(d['graph'], d['method'], 'same_edges', d['same_edges'], 'rel_fval_diff %.1e'%d['rel_fval_diff'], 'max_x_diff %.1e'%d['max_x_diff'], 'speedup run %.2f per-callback %.2f'%(d['speedup_whole_run'], d['speedup_per_callback']))
    else: print('   ', d['impl'], 'time %.2f'%d['time_s'], 'iters', d['iterations'], 'callbacks', d['callbacks'], d['hessian_callbacks'], 'ms/callback %.1f'%d['ms_per_callback'], 'fval', d['fval'])
PY
tail -3 gpurun_out/replay_weighted_hessian.err
