mkdir -p gpurun_out
for st in 1 2 4; do for inf in 2 3; do
timeout 400 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --streams $st --in-flight $inf > gpurun_out/bench28_${st}_${inf}.json 2> gpurun_out/bench28.err
python - <<PY
import json; d=json.load(open("gpurun_out/bench28_${st}_${inf}.json")); print("streams=$st in_flight=$inf", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"])
PY
done; done
