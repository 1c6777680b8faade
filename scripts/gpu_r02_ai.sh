#!/bin/bash
# round 2, call AI: delta entries as a template flavour of the SpMM: tests that edit edges, SpMM timings, bench
mkdir -p gpurun_out
python -m pytest tests/test_gpu_spmm.py tests/test_gpu_krylov.py tests/test_gpu_replicas.py tests/test_gpu_screen.py tests/test_golden.py tests/test_gpu_replay.py tests/test_mex_gateway.py -m gpu -q --timeout=900 2>&1 | tail -3 | cut -c1-300
python scripts/exp_spmm.py new=libkrylov_b200.so > gpurun_out/r02ai_spmm.jsonl 2>&1; cat gpurun_out/r02ai_spmm.jsonl | cut -c1-300
KR_BENCH_EDGES=0 KR_BENCH_C4=0 python bench.py > gpurun_out/r02ai_bench_c3_only.json 2> gpurun_out/r02ai_bench.err; tail -2 gpurun_out/r02ai_bench.err; python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/r02ai_bench_c3_only.json") if l.startswith("{")][-1])
print("value", round(d["value"]), "ms/step", d["ms_per_step"], "e2e", round(d["e2e"]["value"]), "spmm ms", d["roofline"]["ms_per_launch"], "frac", d["roofline"]["frac"], d["clocks"])
PY
