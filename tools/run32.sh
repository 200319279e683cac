mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py -m gpu -x -q --no-header -p no:cacheprovider -k "dwconv or model or oracle or golden or parity" > gpurun_out/t32.log 2>&1; echo "exit $?" >> gpurun_out/t32.log
tail -4 gpurun_out/t32.log
for tma in 1 0; do
MAFB200_DW_TMA=$tma timeout 400 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/bench32_$tma.json 2> gpurun_out/bench32.err
python - <<PY
import json; d=json.load(open("gpurun_out/bench32_$tma.json")); print("dw_tma=$tma", d["value"], d["ms_per_step"], d["breakdown_ms"], "e2e", d["e2e"]["value"]); print({k:v["us_per_forward"] for k,v in d["roofline"]["families"].items()})
PY
done
