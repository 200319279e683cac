mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_preprocess_gpu.py tests/test_postprocess_gpu.py -m gpu -x -q --no-header -p no:cacheprovider > gpurun_out/t31.log 2>&1; echo "exit $?" >> gpurun_out/t31.log
tail -30 gpurun_out/t31.log
