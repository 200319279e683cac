#!/bin/bash
# Round-2 profiling call: ncu launch lists (forward N bs32; loss bs16), --set full of the depth-wise k>=7 kernels, S / M bench lines.
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv \
  --log-file gpurun_out/r2k_launches.csv python tools/one_forward.py > gpurun_out/r2k_ncu.log 2>&1; tail -1 gpurun_out/r2k_ncu.log
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 100 --csv \
  --log-file gpurun_out/r2k_launches_loss.csv python tools/one_loss.py > gpurun_out/r2k_ncu_loss.log 2>&1; tail -1 gpurun_out/r2k_ncu_loss.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:dwconv_kernel -s 10 -c 10 -o gpurun_out/r2k_dwconv_full \
  python tools/one_forward.py > gpurun_out/r2k_ncu_full.log 2>&1; tail -1 gpurun_out/r2k_ncu_full.log
timeout 600 python bench.py --variant s --batch 64 --no-cpu-baseline --steps 100 > gpurun_out/r2k_bench_s_bs64.json 2> gpurun_out/r2k_bench_s.err; echo "s exit $?"
timeout 600 python bench.py --variant m --batch 32 --no-cpu-baseline --steps 100 > gpurun_out/r2k_bench_m_bs32.json 2> gpurun_out/r2k_bench_m.err; echo "m exit $?"
python - <<'PY'
import json
for n in ("s_bs64", "m_bs32"):
    try:
        d = json.load(open(f"gpurun_out/r2k_bench_{n}.json"))
        print(n, d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"], "lat", d["latency_ms_per_batch"]["p50"],
              {k: round(v["us_per_forward"]) for k, v in d["roofline"]["families"].items()}, d["whole_step"]["achieved_bw_frac"], d["whole_step"]["achieved_flops_frac"], d["clocks"])
    except Exception as e:
        print(n, "FAILED", e)
PY
