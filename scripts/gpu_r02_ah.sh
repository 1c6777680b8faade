#!/bin/bash
mkdir -p gpurun_out
python scripts/exp_spmm.py new=libkrylov_b200.so nodelta=libkrylov_nodelta.so new_again=libkrylov_b200.so nodelta_again=libkrylov_nodelta.so > gpurun_out/r02ah_spmm_variants_nodelta.jsonl 2>&1; cat gpurun_out/r02ah_spmm_variants_nodelta.jsonl | cut -c1-330
