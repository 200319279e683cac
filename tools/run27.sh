mkdir -p gpurun_out
timeout 400 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --variant s --batch 64 > gpurun_out/bench27_s.json 2> gpurun_out/bench27_s.err; tail -2 gpurun_out/bench27_s.err
timeout 400 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --variant m --batch 32 > gpurun_out/bench27_m.json 2> gpurun_out/bench27_m.err; tail -2 gpurun_out/bench27_m.err
python - <<PY
import json
for v in "sm":
    d=json.load(open(f"gpurun_out/bench27_{v}.json")); print(v, d["value"], d["ms_per_step"], d["breakdown_ms"], "e2e", d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["kernel"][:30])
PY
