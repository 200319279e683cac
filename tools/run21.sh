mkdir -p gpurun_out
for tc in 0 1; do for st in 1 4; do
MAFB200_DW_TC=$tc timeout 400 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --streams $st > gpurun_out/bench21_${tc}_${st}.json 2> gpurun_out/bench21.err
python - <<PY
import json; d=json.load(open("gpurun_out/bench21_${tc}_${st}.json")); print("tc=$tc streams=$st", d["value"], d["ms_per_step"], d["breakdown_ms"]); print({k:v["us_per_forward"] for k,v in d["roofline"]["families"].items()})
PY
done; done
