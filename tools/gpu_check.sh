#!/bin/bash
# One GPU-box call that re-validates the whole path: GPU parity tests, smoke, the default bench line and the ncu
# launch list of the same command family.  Usage (from the repo root, through gpurun):
#   gpurun --timeout 1800 -- 'bash tools/gpu_check.sh'
# Outputs land in gpurun_out/ (scratch); copy what should be judged into profiles/.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --no-header -p no:cacheprovider > gpurun_out/gpu_tests.log 2>&1; echo "exit $?" >> gpurun_out/gpu_tests.log
tail -4 gpurun_out/gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "exit $?" >> gpurun_out/bench_default.err
tail -2 gpurun_out/bench_default.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_default.json"))
print(d["value"], d["ms_per_step"], d["breakdown_ms"], "e2e", d["e2e"]["value"], "latency p50", d["latency_ms_per_batch"]["p50"],
      "roofline", d["roofline"]["frac"], "cpu", d["cpu_baseline"]["value"] if d.get("cpu_baseline") else None, d["clocks"])
PY
if [ "$1" == "--profile" ]; then
  timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/launches.csv python tools/one_forward.py > gpurun_out/ncu_launches.log 2>&1
  tail -1 gpurun_out/ncu_launches.log
fi
