mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py tests/test_preprocess_gpu.py -m gpu -x -q --no-header -p no:cacheprovider -k "stem or dtypes or batch_feeds" > gpurun_out/t36.log 2>&1; echo "exit $?" >> gpurun_out/t36.log
tail -4 gpurun_out/t36.log
timeout 500 python tools/exp_e2e.py 2>&1 | tail -3
