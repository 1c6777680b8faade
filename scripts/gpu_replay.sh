#!/bin/bash
mkdir -p gpurun_out
python scripts/replay_unweighted.py --graphs oregon_A0 --oracle > gpurun_out/replay_A0_oracle.jsonl 2> gpurun_out/replay_err.log
python scripts/replay_unweighted.py --graphs oregon_A1,oregon_A8,transport_Rome,transport_Barcelona > gpurun_out/replay_more.jsonl 2>> gpurun_out/replay_err.log
cat gpurun_out/replay_A0_oracle.jsonl gpurun_out/replay_more.jsonl | cut -c1-420; tail -3 gpurun_out/replay_err.log
