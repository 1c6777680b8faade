#!/bin/bash
mkdir -p gpurun_out
./scripts/l2_gather > gpurun_out/l2_gather.json 2> gpurun_out/l2_gather.err; cat gpurun_out/l2_gather.json
timeout 1200 python scripts/exp_spmm.py \
  base=libkrylov_b200.so \
  rowsort=libkrylov_b200.so,KR_SPMM_ROWSORT=1 \
  hub1_1k=libkrylov_b200_hub1.so,KR_SPMM_HUBS=1024 \
  hub2_1k=libkrylov_b200_hub2.so,KR_SPMM_HUBS=1024 \
  hub3_1k=libkrylov_b200_hub3.so,KR_SPMM_HUBS=1024 \
  hub2_512=libkrylov_b200_hub2.so,KR_SPMM_HUBS=512 \
  hub2_4k=libkrylov_b200_hub2.so,KR_SPMM_HUBS=4096 \
  hub2_1k_rowsort=libkrylov_b200_hub2.so,KR_SPMM_HUBS=1024,KR_SPMM_ROWSORT=1 \
  > gpurun_out/exp_h.jsonl 2> gpurun_out/exp_h.err
cat gpurun_out/exp_h.jsonl; tail -3 gpurun_out/exp_h.err
