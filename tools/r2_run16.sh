#!/bin/bash
# dwpw: A-tile rows = 25 * warp + pixel (immediate-offset stores): parity in both tile shapes, per-layer times, whole path
for v in 1 0; do
  echo "SMALL=$v: $(MAFB200_DWPW_SMALL=$v timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_blocks_gpu.py -m gpu -q -x --no-header -p no:cacheprovider -k 'dwconv_conv1x1 or block' 2>&1 | tail -1)"
  echo "SMALL=$v: $(MAFB200_DWPW_SMALL=$v python tools/bench_dwpw.py)"
done
bash tools/ab_bench.sh rows1 rows2
