mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_model_gpu.py -m gpu -q --no-header -p no:cacheprovider -s > gpurun_out/t3.log 2>&1; echo "exit $?" >> gpurun_out/t3.log
timeout 200 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "exit $?" >> gpurun_out/smoke.log
timeout 400 python bench.py --steps 50 --warmup 5 > gpurun_out/bench3.json 2> gpurun_out/bench3.err; echo "exit $?" >> gpurun_out/bench3.err
timeout 300 python bench.py --steps 20 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/bench3_eager.json 2>> gpurun_out/bench3.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1a.csv python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
grep -E "passed|failed|vs oracle|vs reference|worst|input" gpurun_out/t3.log | tail; cat gpurun_out/smoke.log | tail -3; cat gpurun_out/bench3.json; tail -3 gpurun_out/bench3.err
