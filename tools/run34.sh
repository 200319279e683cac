mkdir -p gpurun_out
for c in 2 1; do for inf in 2 3; do
MAFB200_GEMM_CTAS_PER_SM=$c timeout 400 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --in-flight $inf > gpurun_out/bench34_${c}_${inf}.json 2> gpurun_out/bench34.err
python - <<PY
import json; d=json.load(open("gpurun_out/bench34_${c}_${inf}.json")); print("ctas/SM=$c in_flight=$inf", d["value"], d["ms_per_step"], d["breakdown_ms"], "e2e", d["e2e"]["value"], d.get("latency_ms_per_batch",{}).get("p50"))
PY
done; done
