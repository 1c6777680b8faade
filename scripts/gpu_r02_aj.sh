#!/bin/bash
mkdir -p gpurun_out
python scripts/exp_spmm.py base=libkrylov_b200.so freeload=libkrylov_freeload.so base_again=libkrylov_b200.so freeload_again=libkrylov_freeload.so > gpurun_out/r02aj_spmm_variants.jsonl 2>&1; cat gpurun_out/r02aj_spmm_variants.jsonl | cut -c1-300
