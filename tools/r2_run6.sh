#!/bin/bash
# Round-2 GPU call #6: dwpw tail channel block by cp.async (C = 72 / 144): parity + A/B.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_blocks_gpu.py -m gpu -q -x --no-header -p no:cacheprovider -k "dwconv or blocks or block" 2>&1 | tail -3
bash tools/ab_bench.sh tail1 tail2
