#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_pairs_small.py tests/test_gpu_replicas.py -m gpu -q --timeout=900 2>&1 | tail -12 | cut -c1-300
python scripts/time_wide_qr.py | tee gpurun_out/r02t_time_wide_qr.jsonl
KR_QR_HOUSEHOLDER=1 python scripts/time_wide_qr.py | tee -a gpurun_out/r02t_time_wide_qr.jsonl
