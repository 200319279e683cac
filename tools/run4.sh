mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider -x > gpurun_out/t4.log 2>&1; echo "exit $?" >> gpurun_out/t4.log
tail -3 gpurun_out/t4.log
timeout 400 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench4.json 2> gpurun_out/bench4.err; echo "exit $?" >> gpurun_out/bench4.err
cat gpurun_out/bench4.json | cut -c1-600
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_r1b.csv python tools/one_forward.py > gpurun_out/ncu4a.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 108 -c 1 -o gpurun_out/prof_gemm_L20m0conv1 -f python tools/one_forward.py > gpurun_out/ncu4b.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:dwconv_kernel -s 22 -c 1 -o gpurun_out/prof_dw_L20m0dw5 -f python tools/one_forward.py > gpurun_out/ncu4c.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:dwconv_kernel -s 19 -c 1 -o gpurun_out/prof_dw_L8m0dw9 -f python tools/one_forward.py > gpurun_out/ncu4d.log 2>&1
ls -la gpurun_out/*.ncu-rep
