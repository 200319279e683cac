timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py -m gpu -x -q --no-header -p no:cacheprovider -k "decode or detect or full_size" 2>&1 | tail -3
for i in 1 2; do
timeout 400 python bench.py --steps 100 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'lat', d['latency_ms_per_batch']['p50'])"
done
