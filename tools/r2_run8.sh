#!/bin/bash
# dwpw with 10x10 tiles / 128-thread CTAs (4-5 CTAs per SM) vs 10x20 / 256 threads: parity, per-layer times, whole path.
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_blocks_gpu.py -m gpu -q -x --no-header -p no:cacheprovider -k "dwconv or block" 2>&1 | tail -3
python tools/bench_dwpw.py
MAFB200_DWPW_SMALL=0 python tools/bench_dwpw.py
bash tools/ab_bench.sh small1 big:MAFB200_DWPW_SMALL=0 small2
