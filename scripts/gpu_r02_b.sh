#!/bin/bash
# round 2, call B: burst + sustained SpMM timing of the lane-width variants (one predicated tail round)
mkdir -p gpurun_out
python -m pytest tests/test_gpu_spmm.py -m gpu -q -x --timeout=900 > gpurun_out/r02b_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02b_pytest.log; tail -2 gpurun_out/r02b_pytest.log
python scripts/exp_spmm.py r01=libkrylov_r01.so cpl4=libkrylov_b200.so cpl2=libkrylov_b200.so,KR_SPMM_CPL=2 \
   cpl4_pf0=libkrylov_pf0.so cpl2_pf0=libkrylov_pf0.so,KR_SPMM_CPL=2 r01_again=libkrylov_r01.so \
   2>&1 | tee gpurun_out/r02b_spmm_variants.jsonl
