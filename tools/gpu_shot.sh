#!/bin/bash
# one gpurun call: parity suites that cover the GRU kernels, then timing + per-kernel table
mkdir -p gpurun_out
run() { name=$1; shift; ( time timeout 600 "$@" ) > gpurun_out/$name.log 2>&1; echo "rc=$?" >> gpurun_out/$name.log; }
run t_par python -m pytest tests/test_forward_parity_gpu.py tests/test_configs_gpu.py tests/test_backward_parity_gpu.py -q -s -x
grep -E "passed|failed|FAILED|Error" gpurun_out/t_par.log | tail -10
run ab_default python tools/ab_time.py prof
grep -h "^\[" gpurun_out/ab_default.log
grep -A8 "eager eval forward" gpurun_out/ab_default.log | cut -c1-60,140-250 | tail -6
