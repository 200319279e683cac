mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/u14.log
./tools/ubench/pipes >> gpurun_out/u14.log 2>&1
cat gpurun_out/u14.log
timeout 400 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench14.json 2> gpurun_out/bench14.err; echo "exit $?" >> gpurun_out/bench14.err
python - <<PY
import json; d=json.load(open("gpurun_out/bench14.json")); print(d["value"], d["ms_per_step"], d["breakdown_ms"], "e2e", d["e2e"]["value"]); print(d["roofline"]["families"])
PY
timeout 900 python -m pytest tests -m gpu -x -q --no-header -p no:cacheprovider > gpurun_out/t14.log 2>&1; echo "exit $?" >> gpurun_out/t14.log
tail -5 gpurun_out/t14.log
