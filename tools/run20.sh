mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q --no-header -p no:cacheprovider -k "dwconv" > gpurun_out/t20.log 2>&1; echo "exit $?" >> gpurun_out/t20.log
tail -4 gpurun_out/t20.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:dwconv -s 16 -c 16 --csv --log-file gpurun_out/dw20.csv python tools/one_forward.py > gpurun_out/ncu20.log 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/dw20.csv")) if len(r)>10]
hdr=rows[0]; tot=0
for r in rows[1:]:
    v=float(r[hdr.index("Metric Value")]); tot+=v
    print(r[hdr.index("Kernel Name")][:30], r[hdr.index("Grid Size")], v/1e3)
print("total us", tot/1e3)
PY
timeout 400 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench20.json 2> gpurun_out/bench20.err; echo "exit $?" >> gpurun_out/bench20.err
python - <<PY
import json; d=json.load(open("gpurun_out/bench20.json")); print(d["value"], d["ms_per_step"], d["breakdown_ms"], "e2e", d["e2e"]["value"]); print(d["roofline"]["families"])
PY
