#!/bin/bash
# round 2, call W (2 GPUs): bench with the screened round, N = 1 and N = 2
mkdir -p gpurun_out
CUDA_VISIBLE_DEVICES=0 python bench.py --steps 3 --warmup 3 > gpurun_out/r02w_bench_1gpu.json 2> gpurun_out/r02w_bench_1gpu.err; echo "bench1 rc $?"; tail -3 gpurun_out/r02w_bench_1gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r02w_bench_2gpu.json 2> gpurun_out/r02w_bench_2gpu.err; echo "bench2 rc $?"; tail -3 gpurun_out/r02w_bench_2gpu.err
python - <<'PY'
import json
for f in ("gpurun_out/r02w_bench_1gpu.json", "gpurun_out/r02w_bench_2gpu.json"):
    lines = [l for l in open(f) if l.startswith("{")]
    d = json.loads(lines[-1])
    s = d["secondary"]
    print(f, "value", round(d["value"]), "C5 exact", round(s["value"]), "screened", s.get("screened_round"))
PY
