#!/bin/bash
# round 2, call E: where the wide-path time goes (ncu launch list of one fun_and_grad call) + new parity test
mkdir -p gpurun_out
python -m pytest tests/test_gpu_krylov.py tests/test_mex_gateway.py -m gpu -q --timeout=900 -k "fun_and_grad or gateway or self_loop" > gpurun_out/r02e_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02e_pytest.log; tail -5 gpurun_out/r02e_pytest.log; grep -E "^E |^FAILED|^ERROR" gpurun_out/r02e_pytest.log | head -20
for w in fun_and_grad tfu_rank2 tfu_set centrality; do KR_PROFILE_WIDE=1 python scripts/one_call.py $w grid_England 2>&1 | tail -4; done
python scripts/one_call.py fun_and_grad grid_England > gpurun_out/r02e_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r02e_launches_fun_and_grad_england.csv python scripts/one_call.py fun_and_grad grid_England > gpurun_out/r02e_ncu.log 2>&1
tail -2 gpurun_out/r02e_ncu.log
