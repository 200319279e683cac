mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dwconv -s 22 -c 1 -o gpurun_out/prof19_dwtc_L20 -f python tools/one_forward.py > gpurun_out/ncu19.log 2>&1
ls -la gpurun_out/prof19*
