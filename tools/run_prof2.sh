mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:dwconv_kernel -s 22 -c 1 -o gpurun_out/prof4_dw5 -f python tools/one_forward.py > gpurun_out/ncuE.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 71 -c 1 -o gpurun_out/prof4_L1 -f python tools/one_forward.py > gpurun_out/ncuF.log 2>&1
ls -la gpurun_out/prof4*
