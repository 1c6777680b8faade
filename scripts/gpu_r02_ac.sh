#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_krylov.py -m gpu -q --timeout=900 -k "basis_invariants" 2>&1 | tail -12 | cut -c1-300
KR_QR_HOUSEHOLDER=1 python -m pytest tests/test_gpu_krylov.py -m gpu -q --timeout=900 -k "basis_invariants" 2>&1 | tail -3 | cut -c1-300
