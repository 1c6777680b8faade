#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_screen.py -m gpu -q --timeout=900 2>&1 | tail -3
KR_SCREEN_EXACT=0 python scripts/bench_screen.py | tee gpurun_out/r02ab_bench_screen_c5.json
