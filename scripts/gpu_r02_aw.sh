#!/bin/bash
# round 2, call AW: fused Arnoldi sweeps (entries.cuh: four sweeps per step instead of nine for the first 16 steps):
# parity tests of everything that uses ArnoldiBatch, then config C2 at full size with and without the fusion
mkdir -p gpurun_out
python -m pytest tests/test_gpu_krylov.py tests/test_gpu_screen.py tests/test_reference_goldens.py tests/test_gpu_differential.py tests/test_gpu_configs.py tests/test_gpu_dropin_matlab.py tests/test_gpu_replay.py -m gpu -q --timeout=900 > gpurun_out/r02aw_pytest.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/r02aw_pytest.log | cut -c1-200; grep -E "^E  |^FAILED|^ERROR" gpurun_out/r02aw_pytest.log | cut -c1-300 | head
for f in 0 1; do
  KR_ARNOLDI_FUSED=$f timeout 600 python scripts/bench_c2.py --graph transport_Vermont --check 12 > gpurun_out/r02aw_c2_fused$f.json 2> gpurun_out/r02aw_c2_fused$f.err; echo "fused=$f rc $?"; cut -c1-700 gpurun_out/r02aw_c2_fused$f.json
done
