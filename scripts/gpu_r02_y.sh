#!/bin/bash
# round 2, call Y (8 GPUs): torchrun bench at N = 8 (weak + strong + C5 split + screened round), replicas in one process
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/r02y_bench_${N}gpu.json 2> gpurun_out/r02y_bench_${N}gpu.err; echo "bench$N rc $?"; tail -3 gpurun_out/r02y_bench_${N}gpu.err
python - <<PY
import json
lines = [l for l in open("gpurun_out/r02y_bench_${N}gpu.json") if l.startswith("{")]
d = json.loads(lines[-1])
s = d["secondary"]
print("N", d["n_gpus"], "value", round(d["value"]), "ms", d["ms_per_step"], "e2e", round(d["e2e"]["value"]), "strong", d["strong_scaling"])
print("C5 exact", round(s["value"]), s["ms_per_round"], "screened", s.get("screened_round"))
PY
KR_BENCH_C5_CAND=32768 python scripts/bench_replicas.py > gpurun_out/r02y_bench_replicas_${N}gpu.json 2> gpurun_out/r02y_bench_replicas.err; cat gpurun_out/r02y_bench_replicas_${N}gpu.json; tail -2 gpurun_out/r02y_bench_replicas.err
