bash tools/gpu_check.sh --profile
timeout 400 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --variant s --batch 64 > gpurun_out/bench_s.json 2>/dev/null
timeout 400 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --variant m --batch 32 > gpurun_out/bench_m.json 2>/dev/null
timeout 400 python bench.py --no-pipeline --no-cpu-baseline > gpurun_out/bench_seq.json 2>/dev/null
python - <<'PY'
import json
for v in ("s","m","seq"):
    d=json.load(open(f"gpurun_out/bench_{v}.json")); print(v, d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "lat", d["latency_ms_per_batch"]["p50"])
PY
