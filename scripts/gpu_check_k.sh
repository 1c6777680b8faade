#!/bin/bash
mkdir -p gpurun_out
timeout 900 python scripts/replay_weighted.py --graphs grid_England,grid_Mexico --oracle > gpurun_out/replay_weighted3.jsonl 2> gpurun_out/replay_weighted3.err
python - <<PY
import json
for l in open('gpurun_out/replay_weighted3.jsonl'):
    d=json.loads(l)
    if 'compare' in d: print(d['graph'], d['method'], 'same_edges', d['same_edges'], 'rel_fval_diff %.1e'%d['rel_fval_diff'], 'speedup run %.2f per-callback %.2f'%(d['speedup_whole_run'], d['speedup_per_callback']))
    else: print('   ', d['impl'], 'time %.2f'%d['time_s'], 'callbacks', d['callbacks'], 'ms/callback %.1f'%d['ms_per_callback'])
PY
tail -3 gpurun_out/replay_weighted3.err
