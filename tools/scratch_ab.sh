mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q --no-header -p no:cacheprovider -k "stem" > gpurun_out/ab_model.log 2>&1; tail -2 gpurun_out/ab_model.log
run() { name=$1; shift; args="$1"; shift; env "$@" timeout 300 python bench.py --no-cpu-baseline --steps 300 $args > gpurun_out/ab_$name.json 2>gpurun_out/ab_$name.err; python -c "
import json; d=json.load(open('gpurun_out/ab_$name.json')); print('$name', d['value'], d['e2e']['value'], d['latency_ms_per_batch']['p50'], {k:v['us_per_forward'] for k,v in d['roofline']['families'].items()})"; }
run base1 "" A=1
run s2 "--streams 2" A=1
run s6 "--streams 6" A=1
run s1 "--streams 1" A=1
run base2 "" A=1
