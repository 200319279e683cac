#!/bin/bash
# Round-2 GPU call #1: parity tests with the new readings, compute-sanitizer passes over the hand-swizzled kernels,
# and the configs nobody but the builder had measured (S bs64, M bs32).  Outputs -> gpurun_out/.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s --no-header -p no:cacheprovider > gpurun_out/r2a_gpu_tests.log 2>&1; echo "exit $?" >> gpurun_out/r2a_gpu_tests.log
grep -a "PARITY\|passed\|failed\|exit" gpurun_out/r2a_gpu_tests.log | tail -60
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python -m pytest tests/test_kernels_gpu.py -q --no-header -p no:cacheprovider \
    -k "dwpw or poolpw or conv1x1 or dwconv_conv1x1 or maxpool2x2_conv1x1" > gpurun_out/r2a_sanitizer_$tool.log 2>&1
  echo "exit $?" >> gpurun_out/r2a_sanitizer_$tool.log
  grep -a "ERROR SUMMARY\|RACECHECK SUMMARY\|passed\|failed\|exit" gpurun_out/r2a_sanitizer_$tool.log | tail -5
done
timeout 600 python bench.py > gpurun_out/r2a_bench_n.json 2> gpurun_out/r2a_bench_n.err; echo "n exit $?"
timeout 600 python bench.py --variant s --batch 64 --no-cpu-baseline --steps 100 > gpurun_out/r2a_bench_s_bs64.json 2> gpurun_out/r2a_bench_s.err; echo "s exit $?"
timeout 600 python bench.py --variant m --batch 32 --no-cpu-baseline --steps 100 > gpurun_out/r2a_bench_m_bs32.json 2> gpurun_out/r2a_bench_m.err; echo "m exit $?"
python - <<'PY'
import json
for n in ("n", "s_bs64", "m_bs32"):
    try:
        d = json.load(open(f"gpurun_out/r2a_bench_{n}.json"))
        print(n, d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"], "p50", d["latency_ms_per_batch"]["p50"],
              "roofline", d["roofline"]["kernel"][:24], d["roofline"]["frac"], d["whole_step"], d["clocks"])
    except Exception as e:
        print(n, "FAILED", e)
PY
