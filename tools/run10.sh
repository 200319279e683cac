mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider > gpurun_out/t20.log 2>&1; echo "exit $?" >> gpurun_out/t20.log
tail -15 gpurun_out/t20.log
