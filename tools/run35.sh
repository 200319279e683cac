mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py -m gpu -x -q --no-header -p no:cacheprovider -k "nms or detect or end_to_end" > gpurun_out/t35.log 2>&1; echo "exit $?" >> gpurun_out/t35.log
tail -12 gpurun_out/t35.log
timeout 400 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --variant s --batch 64 > gpurun_out/bench35_s.json 2> gpurun_out/bench35_s.err; tail -2 gpurun_out/bench35_s.err
timeout 400 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --variant m --batch 32 > gpurun_out/bench35_m.json 2> gpurun_out/bench35_m.err; tail -2 gpurun_out/bench35_m.err
timeout 400 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/bench35_n.json 2> gpurun_out/bench35_n.err; tail -2 gpurun_out/bench35_n.err
python - <<PY
import json
for v in "smn":
    d=json.load(open(f"gpurun_out/bench35_{v}.json")); print(v, d["value"], d["ms_per_step"], d["breakdown_ms"], "e2e", d["e2e"]["value"], d["latency_ms_per_batch"]["p50"])
PY
