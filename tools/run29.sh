mkdir -p gpurun_out
# launch list of the same command family as the bench (one eager forward + NMS after a warm-up pass)
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1d.csv python tools/one_forward.py > gpurun_out/ncu29.log 2>&1
tail -2 gpurun_out/ncu29.log
# full capture of the dominant kernel family: L20.conv1 (3-source fusion GEMM, 80x80) and L20.m0.conv1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 107 -c 2 -o gpurun_out/prof29_gemm_L20 -f python tools/one_forward.py > gpurun_out/ncu29b.log 2>&1
# bench lines: default (pipelined, with cpu baseline), sequential, reference arm
timeout 900 python bench.py > gpurun_out/bench29_default.json 2> gpurun_out/bench29_default.err; echo "exit $?" >> gpurun_out/bench29_default.err
timeout 600 python bench.py --no-pipeline --no-cpu-baseline > gpurun_out/bench29_sequential.json 2> gpurun_out/bench29_seq.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench29_reference.json 2> gpurun_out/bench29_ref.err
cat gpurun_out/bench29_default.json | head -c 3000; echo; cat gpurun_out/bench29_reference.json
