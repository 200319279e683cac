mkdir -p gpurun_out
timeout 400 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench14_n.json 2> gpurun_out/bench14.err; echo "exit n $?" >> gpurun_out/bench14.err
timeout 400 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --variant s --batch 64 > gpurun_out/bench14_s.json 2>> gpurun_out/bench14.err; echo "exit s $?" >> gpurun_out/bench14.err
timeout 400 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --variant m --batch 32 > gpurun_out/bench14_m.json 2>> gpurun_out/bench14.err; echo "exit m $?" >> gpurun_out/bench14.err
for v in n s m; do python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench14_$v.json")); print("$v", d["value"], d["ms_per_step"], d["breakdown_ms"], "e2e", d["e2e"]["value"], d["roofline"]["families"])
except Exception as e: print("$v failed", e)
PY
done
grep -v "^$" gpurun_out/bench14.err | tail -5
