mkdir -p gpurun_out
for st in 1 4; do
timeout 400 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --streams $st > gpurun_out/bench13_s$st.json 2> gpurun_out/bench13.err; echo "exit $?" >> gpurun_out/bench13.err
python - <<PY
import json; d=json.load(open("gpurun_out/bench13_s$st.json")); print("streams $st", d["value"], d["ms_per_step"], d["breakdown_ms"], d["e2e"]["value"])
PY
done
tail -2 gpurun_out/bench13.err
