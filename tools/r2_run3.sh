#!/bin/bash
# Round-2 GPU call #3: K4 (fused DepthBottleneckUni) — kernel test under a hard timeout (a pipeline bug must not hang
# the box), sanitizer pass, whole suite, A/B bench.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -x --no-header -p no:cacheprovider -k "bottleneck_fused" > gpurun_out/r2c_k4_kernel.log 2>&1; rc=$?
echo "k4 kernel test exit $rc"; tail -25 gpurun_out/r2c_k4_kernel.log | cut -c1-300
if [ $rc -ne 0 ]; then exit 0; fi
timeout 600 compute-sanitizer --tool racecheck --print-limit 10 python -m pytest tests/test_kernels_gpu.py -q --no-header -p no:cacheprovider -k "bottleneck_fused and not 160" > gpurun_out/r2c_sanitizer_racecheck_k4.log 2>&1
grep -a "RACECHECK SUMMARY\|passed\|failed" gpurun_out/r2c_sanitizer_racecheck_k4.log | tail -3
timeout 600 compute-sanitizer --tool memcheck --print-limit 10 python -m pytest tests/test_kernels_gpu.py -q --no-header -p no:cacheprovider -k "bottleneck_fused and not 160" > gpurun_out/r2c_sanitizer_memcheck_k4.log 2>&1
grep -a "ERROR SUMMARY\|passed\|failed" gpurun_out/r2c_sanitizer_memcheck_k4.log | tail -3
timeout 1500 python -m pytest tests -m gpu -q -s --no-header -p no:cacheprovider > gpurun_out/r2c_gpu_tests.log 2>&1; echo "exit $?" >> gpurun_out/r2c_gpu_tests.log
grep -a "PARITY\|passed\|failed\|exit\|^FAILED\|^E  " gpurun_out/r2c_gpu_tests.log | cut -c1-330 | tail -40
bash tools/ab_bench.sh k4on k4off:MAFB200_BNECK=0 k4on2
