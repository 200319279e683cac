#!/bin/bash
# Round-2 GPU call #5: depth-wise k = 7 / 9 with the tap nest expanded at compile time (A/B against HEAD~ numbers in
# profiles/r02_e_*): kernel parity, then images/s + per-family us for N and M, 3 vs 4 CTAs/SM for k = 9.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x --no-header -p no:cacheprovider -k "dwconv" 2>&1 | tail -3
bash tools/ab_bench.sh base1 dw9b4:MAFB200_LIB=maf_yolo_b200/libmafb200_dw9b4.so cb64k7:MAFB200_DW_CB64_MAXK=7 base2 \
  "m_base::--variant m --batch 32 --steps 100" "m_dw9b4:MAFB200_LIB=maf_yolo_b200/libmafb200_dw9b4.so:--variant m --batch 32 --steps 100" \
  "m_cb64k7:MAFB200_DW_CB64_MAXK=7:--variant m --batch 32 --steps 100"
