mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider -x > gpurun_out/t8.log 2>&1; echo "exit $?" >> gpurun_out/t8.log
tail -3 gpurun_out/t8.log
timeout 400 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench8.json 2> gpurun_out/bench8.err; echo "exit $?" >> gpurun_out/bench8.err
cat gpurun_out/bench8.json | cut -c1-300; tail -2 gpurun_out/bench8.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_r1f.csv python tools/one_forward.py > gpurun_out/ncu8a.log 2>&1
