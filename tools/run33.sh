mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dwconv_kernel -s 22 -c 1 -o gpurun_out/prof33_dw_L20 -f python tools/one_forward.py > gpurun_out/ncu33.log 2>&1
ls -la gpurun_out/prof33*
