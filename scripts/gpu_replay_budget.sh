#!/bin/bash
# budget sweeps of Tests/test_unweighted_break_budget.m / _make_budget.m (Figures 1-4): k = 10..100, Q in {50, 250, 1000}
# on the road networks we carry as fixtures; the oracle rides along on the cheap corner only.
mkdir -p gpurun_out
out=gpurun_out/replay_budget.jsonl; : > $out
python scripts/replay_unweighted.py --graphs transport_Anaheim --k 10 --Q 50 --oracle >> $out 2>> gpurun_out/replay_budget.err
for Q in 50 250 1000; do for k in 10 50 100; do
  timeout 300 python scripts/replay_unweighted.py --graphs transport_Anaheim,transport_Rome,transport_Barcelona --k $k --Q $Q >> $out 2>> gpurun_out/replay_budget.err
done; done
python - <<PY
import json
for l in open('$out'):
    d=json.loads(l)
    if d.get('method','').startswith('GREEDY'):
        print(d['graph'], d['method'], 'k',d['k'],'Q',d['Q'], 'time %.3f'%d['time_s'], 'edges/s %.0f'%d['edges_per_s'], 'tr_var %.4g'%d['tr_variation'], ('same_edges %s oracle %.2fs'%(d['same_edges'], d['oracle_time_s'])) if 'same_edges' in d else '')
PY
tail -2 gpurun_out/replay_budget.err
