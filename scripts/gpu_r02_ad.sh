#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_krylov.py -m gpu -q --timeout=900 -k "basis_invariants" 2>&1 | tail -6 | cut -c1-300
KR_QR_HOUSEHOLDER=1 python -m pytest tests/test_gpu_krylov.py -m gpu -q --timeout=900 -k "basis_invariants" 2>&1 | tail -3 | cut -c1-300
python -m pytest tests/test_gpu_screen.py -m gpu -q --timeout=900 2>&1 | tail -6 | cut -c1-300
python scripts/diag_screen_shards.py 2>&1 | tail -5
