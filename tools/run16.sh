mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 73 -c 2 -o gpurun_out/prof16_gemm_L2m0 -f python tools/one_forward.py > gpurun_out/ncu16a.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 127 -c 1 -o gpurun_out/prof16_gemm_L31stem -f python tools/one_forward.py > gpurun_out/ncu16b.log 2>&1
tail -3 gpurun_out/ncu16a.log gpurun_out/ncu16b.log
ls -la gpurun_out/prof16*
