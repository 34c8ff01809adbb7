#!/bin/bash
# ncu --set full of the spectral-block kernels (second forward: warm), one launch each
mkdir -p gpurun_out
# 17 launches per forward; skip the first forward, then capture the first tc3 (gft), first chain, first tc3 (head)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'tc3_kernel|glu_chain_h_kernel' -s 6 -c 3 \
  -o gpurun_out/r02_spec -f python tools/ncu_fwd.py 3 > gpurun_out/ncu_spec.log 2>&1
tail -5 gpurun_out/ncu_spec.log
ls -la gpurun_out/*.ncu-rep
