mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --no-header -p no:cacheprovider > gpurun_out/t26.log 2>&1; echo "exit $?" >> gpurun_out/t26.log
tail -5 gpurun_out/t26.log
timeout 400 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench26.json 2> gpurun_out/bench26.err; echo "exit $?" >> gpurun_out/bench26.err
tail -3 gpurun_out/bench26.err
python - <<PY
import json; d=json.load(open("gpurun_out/bench26.json")); print(d["value"], d["ms_per_step"], d["breakdown_ms"], "e2e", d["e2e"])
PY
timeout 400 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-pipeline > gpurun_out/bench26b.json 2> gpurun_out/bench26b.err
python - <<PY
import json; d=json.load(open("gpurun_out/bench26b.json")); print("sequential", d["value"], d["ms_per_step"], d["breakdown_ms"], "e2e", d["e2e"]["value"])
PY
