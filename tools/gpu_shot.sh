#!/bin/bash
# one gpurun call: unit tests of the new kernels, then A/B timings (one process per switch setting)
mkdir -p gpurun_out
run() { name=$1; shift; ( time timeout 600 "$@" ) > gpurun_out/$name.log 2>&1; echo "rc=$?" >> gpurun_out/$name.log; }
run t_spec python -m pytest tests/test_spec_tc_gpu.py tests/test_ddp_global_graph_gpu.py -q -s
grep -E "passed|failed|max\|err|ddp_global" gpurun_out/t_spec.log | tail -30
run t_par python -m pytest "tests/test_forward_parity_gpu.py" tests/test_glu_tc_gpu.py tests/test_backward_parity_gpu.py tests/test_trainer_gpu.py -q -s
grep -E "passed|failed|FAILED|worst relative" gpurun_out/t_par.log | tail -30
run ab_default python tools/ab_time.py prof
STEMGNN_NO_GFT_TC=1 STEMGNN_NO_FUSED_HEAD=1 run ab_oldspec python tools/ab_time.py
STEMGNN_GLU_NO_MULTICAST=1 run ab_nomc python tools/ab_time.py
STEMGNN_TC_NOSPLIT=1 run ab_nosplit python tools/ab_time.py
grep -h "^\[" gpurun_out/ab_*.log
