run() { name=$1; shift; args="$1"; shift; env "$@" timeout 300 python bench.py --no-cpu-baseline $args > gpurun_out/ab_$name.json 2>gpurun_out/ab_$name.err; python -c "
import json; d=json.load(open('gpurun_out/ab_$name.json')); print('$name', d['value'], d['e2e']['value'], d['latency_ms_per_batch']['p50'], {k:v['us_per_forward'] for k,v in d['roofline']['families'].items()})"; }
run s_base "--variant s --batch 64" A=1
run s_nopool "--variant s --batch 64" MAFB200_POOLPW=0
run s_nopad "--variant s --batch 64" MAFB200_PAD_FILL=0
run m_base "--variant m --batch 32" A=1
run m_nopool "--variant m --batch 32" MAFB200_POOLPW=0
run m_nopad "--variant m --batch 32" MAFB200_PAD_FILL=0
