mkdir -p gpurun_out
for cfg in "1 4" "1 1" "1 2" "1 6" "0 4" "0 1"; do
set -- $cfg
MAFB200_PDL=$1 timeout 200 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --streams $2 > gpurun_out/bench_x.json 2> gpurun_out/bench_x.err; echo "exit $?" >> gpurun_out/bench_x.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_x.json")); print("pdl=$1 streams=$2", d["value"], d["ms_per_step"], d["breakdown_ms"], "e2e", d["e2e"]["value"])
except Exception as e: print("$cfg", "failed", e)
PY
done
