#!/bin/bash
# final evidence: launch list of eager forwards (durations) + ncu --set full of the spectral-block kernels at HEAD
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_forward_final.csv \
  python tools/ncu_fwd.py 3 > gpurun_out/ncu_list.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'tc3_kernel|glu_chain_h_kernel' -s 6 -c 3 \
  -o gpurun_out/r02_spec_v2 -f python tools/ncu_fwd.py 3 > gpurun_out/ncu_spec2.log 2>&1
tail -3 gpurun_out/ncu_spec2.log; ls -la gpurun_out/*.ncu-rep gpurun_out/r02_launches_forward_final.csv
