#!/bin/bash
# Round-2 GPU call #2: K7 (DFL / sigmoid / candidate filter in the head GEMM epilogues) — kernel test, whole GPU suite,
# and the attribution A/Bs of the parity error: K7 on/off (fp16 rounding of the head logits), tanh.approx vs exact SiLU.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -s --no-header -p no:cacheprovider -k "head_pred" > gpurun_out/r2b_k7_kernel.log 2>&1; echo "k7 kernel test exit $?"
tail -15 gpurun_out/r2b_k7_kernel.log | cut -c1-400
timeout 1500 python -m pytest tests -m gpu -q -s --no-header -p no:cacheprovider > gpurun_out/r2b_gpu_tests.log 2>&1; echo "exit $?" >> gpurun_out/r2b_gpu_tests.log
grep -a "PARITY\|passed\|failed\|exit\|^FAILED\|^E  " gpurun_out/r2b_gpu_tests.log | cut -c1-330 | tail -50
for v in "k7off:MAFB200_K7=0" "siluexact:MAFB200_LIB=$PWD/maf_yolo_b200/libmafb200_silu_exact.so" "siluexact_k7off:MAFB200_K7=0,MAFB200_LIB=$PWD/maf_yolo_b200/libmafb200_silu_exact.so"; do
  IFS=':' read -r name envs <<< "$v"
  envs=${envs//,/ }
  env $envs timeout 900 python -m pytest tests/test_model_gpu.py -q -s --no-header -p no:cacheprovider -k "forward_matches_oracle or conditioned or full_size_configs" > gpurun_out/r2b_parity_$name.log 2>&1
  echo "== $name exit $?"; grep -a "PARITY\|passed\|failed" gpurun_out/r2b_parity_$name.log | cut -c1-330
done
for v in "k7on:" "k7off:MAFB200_K7=0" "siluexact:MAFB200_LIB=$PWD/maf_yolo_b200/libmafb200_silu_exact.so"; do
  IFS=':' read -r name envs <<< "$v"
  env $envs timeout 300 python bench.py --no-cpu-baseline --steps 300 > gpurun_out/r2b_bench_$name.json 2> gpurun_out/r2b_bench_$name.err
  python - "$name" <<'PY'
import json, sys
name = sys.argv[1]
try:
    d = json.load(open(f"gpurun_out/r2b_bench_{name}.json"))
    fam = {k: v["us_per_forward"] for k, v in d["roofline"]["families"].items()}
    print(name, d["value"], "e2e", d["e2e"]["value"], "p50 ms", d["latency_ms_per_batch"]["p50"], fam)
except Exception as e:  # noqa: BLE001
    print(name, "FAILED:", e, open(f"gpurun_out/r2b_bench_{name}.err").read()[-600:])
PY
done
