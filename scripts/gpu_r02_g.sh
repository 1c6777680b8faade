#!/bin/bash
mkdir -p gpurun_out
python scripts/diag_wide_R.py grid_England sinh 2>&1 | tail -12
python scripts/diag_wide_R.py transport_Rome cosh 2>&1 | tail -12
echo "--- round-1 library (cuSOLVER QR)"
KR_B200_LIB=$PWD/krylov_robustness_b200/libkrylov_r01.so python scripts/diag_wide.py grid_England cosh sinh 2>&1 | sed -n 1,8p
