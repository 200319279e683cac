mkdir -p gpurun_out
cp maf_yolo_b200/libmafb200.so /tmp/lib_orig.so
for v in orig hint_200 hint_20000; do
if [ $v != orig ]; then cp gpurun_tmp/lib_$v.so maf_yolo_b200/libmafb200.so; else cp /tmp/lib_orig.so maf_yolo_b200/libmafb200.so; fi
timeout 400 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench23_$v.json 2> gpurun_out/bench23.err
python - <<PY
import json; d=json.load(open("gpurun_out/bench23_$v.json")); print("$v", d["value"], d["ms_per_step"], d["breakdown_ms"]); print({k:v["us_per_forward"] for k,v in d["roofline"]["families"].items()})
PY
done
cp /tmp/lib_orig.so maf_yolo_b200/libmafb200.so
