#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; ( time timeout 600 "$@" ) > gpurun_out/$name.log 2>&1; echo "rc=$?" >> gpurun_out/$name.log; }
run t_par python -m pytest tests/test_spec_tc_gpu.py tests/test_glu_tc_gpu.py tests/test_forward_parity_gpu.py tests/test_backward_parity_gpu.py tests/test_trainer_gpu.py -q -x
grep -E "passed|failed|FAILED|Error" gpurun_out/t_par.log | tail -10
run ab_default python tools/ab_time.py prof
STEMGNN_TC_NOSPLIT=1 run ab_nosplit python tools/ab_time.py
grep -h "^\[" gpurun_out/ab_*.log
grep -A12 "eager eval forward" gpurun_out/ab_default.log | cut -c1-60,140-250 | tail -10
grep -A8 "eager train step" gpurun_out/ab_default.log | cut -c1-60,140-250 | tail -6
