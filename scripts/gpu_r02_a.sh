#!/bin/bash
# round 2, call A: correctness of the re-written SpMM (128-/256-bit lanes, unpredicated rounds) + variant timing
mkdir -p gpurun_out
python -m pytest tests/test_gpu_spmm.py tests/test_gpu_expmv.py -m gpu -q -x --timeout=900 > gpurun_out/r02a_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02a_pytest.log; tail -3 gpurun_out/r02a_pytest.log
KR_SPMM_CPL=2 python -m pytest tests/test_gpu_spmm.py -m gpu -q -x --timeout=900 > gpurun_out/r02a_pytest_cpl2.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02a_pytest_cpl2.log; tail -2 gpurun_out/r02a_pytest_cpl2.log
python scripts/exp_spmm.py r01=libkrylov_r01.so cpl4=libkrylov_b200.so cpl2=libkrylov_b200.so,KR_SPMM_CPL=2 \
   cpl4_3cta_u6=libkrylov_v3u6.so cpl4_3cta_u8=libkrylov_v3u8.so cpl2_3cta=libkrylov_v3u8.so,KR_SPMM_CPL=2 r01_again=libkrylov_r01.so cpl4_again=libkrylov_b200.so \
   2>&1 | tee gpurun_out/r02a_spmm_variants.jsonl
