mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_model_gpu.py -m gpu -q --no-header -p no:cacheprovider -s > gpurun_out/t27.log 2>&1; echo "exit $?" >> gpurun_out/t27.log
grep -E "vs oracle|vs reference|worst|passed|failed" gpurun_out/t27.log
timeout 400 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench27.json 2> gpurun_out/bench27.err; echo "exit $?" >> gpurun_out/bench27.err
python - <<PY
import json; d=json.load(open("gpurun_out/bench27.json")); print(d["value"], d["ms_per_step"], d["breakdown_ms"], "e2e", d["e2e"]["value"]); print(d["roofline"]["families"])
PY
