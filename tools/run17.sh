mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q --no-header -p no:cacheprovider -k "dwconv" > gpurun_out/t17.log 2>&1; echo "exit $?" >> gpurun_out/t17.log
tail -25 gpurun_out/t17.log
