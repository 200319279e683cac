mkdir -p gpurun_out
for cap in 0 2 3 4 6; do
MAFB200_GEMM_MAX_STAGES=$cap timeout 400 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench22_${cap}.json 2> gpurun_out/bench22.err
python - <<PY
import json; d=json.load(open("gpurun_out/bench22_${cap}.json")); print("cap=$cap", d["value"], d["ms_per_step"], d["breakdown_ms"]); print({k:v["us_per_forward"] for k,v in d["roofline"]["families"].items()})
PY
done
