#!/bin/bash
# Round-2 GPU call #4: whole GPU suite after K7 / gather / full-graph changes + the default bench line for N, S bs64, M bs32.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s -x --no-header -p no:cacheprovider > gpurun_out/r2e_gpu_tests.log 2>&1; echo "exit $?" >> gpurun_out/r2e_gpu_tests.log
grep -a "passed\|failed\|exit\|^FAILED\|^E  " gpurun_out/r2e_gpu_tests.log | cut -c1-300 | tail -12
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | cut -c1-400
timeout 900 python bench.py > gpurun_out/r2e_bench_n.json 2> gpurun_out/r2e_bench_n.err; echo "n exit $?"; tail -3 gpurun_out/r2e_bench_n.err
timeout 600 python bench.py --variant s --batch 64 --no-cpu-baseline --steps 100 > gpurun_out/r2e_bench_s_bs64.json 2> gpurun_out/r2e_bench_s.err; echo "s exit $?"
timeout 600 python bench.py --variant m --batch 32 --no-cpu-baseline --steps 100 > gpurun_out/r2e_bench_m_bs32.json 2> gpurun_out/r2e_bench_m.err; echo "m exit $?"
python - <<'PY'
import json
for n in ("n", "s_bs64", "m_bs32"):
    try:
        d = json.load(open(f"gpurun_out/r2e_bench_{n}.json"))
        print(n, d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"], "lat", d["latency_ms_per_batch"]["p50"], d["latency_ms_per_batch"].get("p50_reference_shaped_calls"),
              "roofline", d["roofline"]["kernel"][:24], d["roofline"]["frac"], {k: d["whole_step"][k] for k in ("hbm_frac_of_peak", "achieved_bw_frac", "achieved_flops_frac")},
              "lib", d.get("gpu_library_baseline"), "cpu", (d.get("cpu_baseline") or {}).get("value"), "host", d["host_enqueue_ms_per_step"], d["clocks"], "launches", d["gpu_launches_per_step"])
    except Exception as e:
        print(n, "FAILED", e)
PY
