mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 72 -c 1 -o gpurun_out/prof2_gemm_L2conv1 -f python tools/one_forward.py > gpurun_out/ncuA.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 108 -c 1 -o gpurun_out/prof2_gemm_L20m0conv1 -f python tools/one_forward.py > gpurun_out/ncuB.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:nms_select_kernel -s 1 -c 1 -o gpurun_out/prof2_nms_select -f python tools/one_forward.py > gpurun_out/ncuC.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:stem_conv_kernel -s 1 -c 1 -o gpurun_out/prof2_stem -f python tools/one_forward.py > gpurun_out/ncuD.log 2>&1
ls -la gpurun_out/prof2*
