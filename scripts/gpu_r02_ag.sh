#!/bin/bash
mkdir -p gpurun_out
python scripts/exp_spmm.py base=libkrylov_b200.so persist1=libkrylov_b200.so,KR_SPMM_PERSIST=1 persist2=libkrylov_b200.so,KR_SPMM_PERSIST=2 persist4=libkrylov_b200.so,KR_SPMM_PERSIST=4 base_again=libkrylov_b200.so > gpurun_out/r02ag_spmm_variants.jsonl 2>&1; cat gpurun_out/r02ag_spmm_variants.jsonl | cut -c1-330
