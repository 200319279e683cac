#!/bin/bash
# A/B of environment switches on one GPU box: each argument is one bench variant,
# "name[:ENV=VAL[,ENV=VAL...]][:extra bench.py flags]".  ~10 s per variant (300 steps, no CPU baseline), so a whole
# comparison costs about a GPU-minute.  Run the same variant twice (base1 / base2) to see the box's noise (~0.5 %).
#   gpurun --timeout 600 -- 'bash tools/ab_bench.sh base1 nopool:MAFB200_POOLPW=0 s3:MAFB200_DWPW_3CTA=0:"--streams 2" base2'
# Prints images/s (device-resident, e2e), p50 latency of one batch and the event-timed us per kernel family.
mkdir -p gpurun_out
for spec in "$@"; do
  IFS=':' read -r name envs flags <<< "$spec"
  envs=${envs//,/ }
  env $envs timeout 300 python bench.py --no-cpu-baseline --steps 300 $flags > gpurun_out/ab_$name.json 2> gpurun_out/ab_$name.err
  python - "$name" <<'PY'
import json, sys
name = sys.argv[1]
try:
    d = json.load(open(f"gpurun_out/ab_{name}.json"))
    fam = {k: v["us_per_forward"] for k, v in d["roofline"]["families"].items()}
    print(name, d["value"], "e2e", d["e2e"]["value"], "p50 ms", d["latency_ms_per_batch"]["p50"], fam)
except Exception as e:  # noqa: BLE001
    print(name, "FAILED:", e, open(f"gpurun_out/ab_{name}.err").read()[-400:])
PY
done
