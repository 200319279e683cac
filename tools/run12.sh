mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider > gpurun_out/t26.log 2>&1; echo "exit $?" >> gpurun_out/t26.log
tail -3 gpurun_out/t26.log
timeout 500 python bench.py --steps 100 --warmup 10 > gpurun_out/bench26_n.json 2> gpurun_out/bench26.err; echo "exit n $?" >> gpurun_out/bench26.err
timeout 300 python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/bench26_ref.json 2>> gpurun_out/bench26.err; echo "exit ref $?" >> gpurun_out/bench26.err
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_r1t.csv python tools/one_forward.py > gpurun_out/ncu26a.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 108 -c 1 -o gpurun_out/prof5_gemm_L20m0conv1 -f python tools/one_forward.py > gpurun_out/ncu26b.log 2>&1
cat gpurun_out/bench26_n.json | cut -c1-300; cat gpurun_out/bench26_ref.json | cut -c1-200; tail -3 gpurun_out/bench26.err
