#!/bin/bash
# compute-sanitizer memcheck over a small slice of the GPU parity tests
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 99 --log-file gpurun_out/sanitizer_memcheck.log \
  python -m pytest tests -m gpu -q -x --timeout=1400 \
  -k "(spmm_matches_scipy and oregon_A0) or unsymmetric or (edges_vs_oracle and oregon_A0) or leaf or self_loop or (entries_vs_oracle and oregon_A0-exp) or (expmv_vs_oracle and oregon_A0) or shift_with_diagonal or (slq_matches and oregon_A0) or (basis_invariants and 2-True) or normAm" \
  > gpurun_out/sanitizer_pytest.log 2>&1
echo "exit $?" >> gpurun_out/sanitizer_pytest.log
tail -5 gpurun_out/sanitizer_pytest.log; tail -5 gpurun_out/sanitizer_memcheck.log
