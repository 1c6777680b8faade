#!/bin/bash
mkdir -p gpurun_out
python scripts/time_wide_qr.py | tee gpurun_out/r02z_time_wide_qr.jsonl
KR_QR_HOUSEHOLDER=1 python scripts/time_wide_qr.py | tee -a gpurun_out/r02z_time_wide_qr.jsonl
KR_QR_HOUSEHOLDER=1 python -m pytest tests/test_gpu_krylov.py tests/test_gpu_expmv.py -m gpu -q --timeout=900 2>&1 | tail -2
python -m pytest tests/test_gpu_krylov.py tests/test_gpu_pairs_small.py tests/test_mex_gateway.py -m gpu -q --timeout=900 2>&1 | tail -2
