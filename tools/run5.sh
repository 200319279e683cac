mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider -x > gpurun_out/t11.log 2>&1; echo "exit $?" >> gpurun_out/t11.log
tail -3 gpurun_out/t11.log
timeout 400 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench11.json 2> gpurun_out/bench11.err; echo "exit $?" >> gpurun_out/bench11.err
cat gpurun_out/bench11.json | cut -c1-300; tail -2 gpurun_out/bench11.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_r1i.csv python tools/one_forward.py > gpurun_out/ncu11a.log 2>&1
