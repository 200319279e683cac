mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider -s > gpurun_out/t24.log 2>&1; echo "exit $?" >> gpurun_out/t24.log
grep -E "passed|failed|vs oracle|vs reference|worst|FAILED" gpurun_out/t24.log | tail -12
timeout 400 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench24.json 2> gpurun_out/bench24.err; echo "exit $?" >> gpurun_out/bench24.err
python - <<PY
import json; d=json.load(open("gpurun_out/bench24.json")); print(d["value"], d["ms_per_step"], d["breakdown_ms"], "e2e", d["e2e"]["value"]); print(d["roofline"]["families"])
PY
tail -2 gpurun_out/bench24.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_r1r.csv python tools/one_forward.py > gpurun_out/ncu24a.log 2>&1
