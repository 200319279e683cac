#!/bin/bash
# End of round 2: whole GPU suite, smoke, bench lines (N with CPU + library + training legs; S bs64; M bs32), ncu launch list + DRAM traffic.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --no-header -p no:cacheprovider > gpurun_out/r2t_gpu_tests.log 2>&1; echo "exit $?" >> gpurun_out/r2t_gpu_tests.log
grep -a "passed\|failed\|exit\|^FAILED\|^E  " gpurun_out/r2t_gpu_tests.log | cut -c1-300 | tail -6
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | cut -c1-300
timeout 900 python bench.py > gpurun_out/r2t_bench_n.json 2> gpurun_out/r2t_bench_n.err; echo "n exit $?"
timeout 600 python bench.py --variant s --batch 64 --no-cpu-baseline --steps 100 > gpurun_out/r2t_bench_s_bs64.json 2> gpurun_out/r2t_bench_s.err; echo "s exit $?"
timeout 600 python bench.py --variant m --batch 32 --no-cpu-baseline --steps 100 > gpurun_out/r2t_bench_m_bs32.json 2> gpurun_out/r2t_bench_m.err; echo "m exit $?"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv \
  --log-file gpurun_out/r2t_launches.csv python tools/one_forward.py > gpurun_out/r2t_ncu.log 2>&1; tail -1 gpurun_out/r2t_ncu.log
python - <<'PY'
import json
for n in ("n", "s_bs64", "m_bs32"):
    try:
        d = json.load(open(f"gpurun_out/r2t_bench_{n}.json"))
        print(n, d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"], "lat", d["latency_ms_per_batch"]["p50"], "roofline", d["roofline"]["frac"],
              {k: round(v["us_per_forward"]) for k, v in d["roofline"]["families"].items()}, d["whole_step"]["achieved_bw_frac"], d["whole_step"]["achieved_flops_frac"],
              "lib", (d.get("gpu_library_baseline") or {}).get("value"), "cpu", (d.get("cpu_baseline") or {}).get("value"), d["clocks"])
    except Exception as e:
        print(n, "FAILED", e)
PY
