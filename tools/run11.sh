mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider > gpurun_out/t25.log 2>&1; echo "exit $?" >> gpurun_out/t25.log
tail -4 gpurun_out/t25.log
timeout 400 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench25.json 2> gpurun_out/bench25.err; echo "exit $?" >> gpurun_out/bench25.err
python - <<PY
import json; d=json.load(open("gpurun_out/bench25.json")); print(d["value"], d["ms_per_step"], d["breakdown_ms"], "e2e", d["e2e"]["value"])
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_r1s.csv python tools/one_forward.py > gpurun_out/ncu25a.log 2>&1
