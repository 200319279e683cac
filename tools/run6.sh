mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider -x > gpurun_out/t12.log 2>&1; echo "exit $?" >> gpurun_out/t12.log
tail -3 gpurun_out/t12.log
timeout 400 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench12.json 2> gpurun_out/bench12.err; echo "exit $?" >> gpurun_out/bench12.err
cat gpurun_out/bench12.json | cut -c1-300; tail -2 gpurun_out/bench12.err
