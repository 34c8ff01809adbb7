#!/bin/bash
# profiles/sass_summary.txt: Blackwell-specific opcodes per kernel family in the built library (cuobjdump -sass, no GPU needed)
# usage: bash tools/sass_summary.sh > profiles/sass_summary.txt
LIB=stemgnn_b200/libstemgnn_b200.so
PAT='UTCHMMA|UTMALDG[.0-9A-Z]*|UTMASTG[.0-9A-Z]*|LDTM[.0-9a-zA-Z]*|STTM[.0-9a-zA-Z]*|UTCBAR[.0-9A-Z]*|STAS[.0-9A-Z]*|UCGABAR_[A-Z]+|ELECT|SYNCS\.[.0-9A-Z]+|FFMA2|LDS\.128|STS\.128|LD\.E\.128|ST\.E\.128|REDG[.0-9A-Z]*'
echo "# opcode counts per kernel family, all template instantiations summed (cuobjdump -sass $LIB at HEAD)"
for k in gru_tc_cluster_kernel gru_step_tc_kernel glu_chain_h_kernel tc3_kernel laplacian_eig_kernel glu_chain_tc_kernel glu_tc_kernel tc_gemm_kernel gru_cluster_kernel gru_bwd_cluster_kernel sgemm_kernel; do
  echo "== $k"
  cuobjdump -sass "$LIB" | awk -v k="$k" '/Function :/ {on = index($0, k) > 0} on' | grep -oE "\b($PAT)\b" | sort | uniq -c | sort -rn | head -14
done
