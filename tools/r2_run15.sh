#!/bin/bash
# compute-sanitizer over the kernels touched at the end of round 2: loss kernels, depth-wise k = 7 / 9, dwpw with 10 x 10 tiles
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_loss_gpu.py -m gpu -q -x --no-header -p no:cacheprovider -k "end_to_end or no_targets or capacity" > gpurun_out/r2m_memcheck_loss.log 2>&1; tail -4 gpurun_out/r2m_memcheck_loss.log
timeout 900 compute-sanitizer --tool racecheck python -m pytest tests/test_loss_gpu.py -m gpu -q -x --no-header -p no:cacheprovider -k "end_to_end and sparse" > gpurun_out/r2m_racecheck_loss.log 2>&1; tail -4 gpurun_out/r2m_racecheck_loss.log
timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_kernels_gpu.py -m gpu -q -x --no-header -p no:cacheprovider -k "dwconv" > gpurun_out/r2m_memcheck_dw.log 2>&1; tail -4 gpurun_out/r2m_memcheck_dw.log
timeout 1200 compute-sanitizer --tool racecheck python -m pytest tests/test_kernels_gpu.py -m gpu -q -x --no-header -p no:cacheprovider -k "dwconv_conv1x1" > gpurun_out/r2m_racecheck_dwpw.log 2>&1; tail -4 gpurun_out/r2m_racecheck_dwpw.log
