#!/bin/bash
# Round-2 GPU call: whole GPU suite + smoke + default bench line (with the training-side leg) after the dwconv / dwpw / loss work.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --no-header -p no:cacheprovider > gpurun_out/r2k_gpu_tests.log 2>&1; echo "exit $?" >> gpurun_out/r2k_gpu_tests.log
grep -a "passed\|failed\|exit\|^FAILED\|^E  " gpurun_out/r2k_gpu_tests.log | cut -c1-300 | tail -12
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4 | cut -c1-400
timeout 900 python bench.py > gpurun_out/r2k_bench_n.json 2> gpurun_out/r2k_bench_n.err; echo "n exit $?"; tail -3 gpurun_out/r2k_bench_n.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2k_bench_n.json"))
print(d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"], "lat", d["latency_ms_per_batch"]["p50"], "roofline", d["roofline"]["frac"],
      "lib", d.get("gpu_library_baseline", {}).get("value"), "cpu", (d.get("cpu_baseline") or {}).get("value"), d["clocks"])
print("training_side", json.dumps(d.get("training_side")))
PY
