#!/bin/bash
timeout 500 python -m pytest tests/test_loss_gpu.py -m gpu -q --no-header -p no:cacheprovider 2>&1 | tail -3
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 100 --csv \
  --log-file gpurun_out/r2l_launches_loss.csv python tools/one_loss.py > gpurun_out/r2l_ncu_loss.log 2>&1; tail -1 gpurun_out/r2l_ncu_loss.log
python - <<'PY'
import torch, sys
sys.path.insert(0, '.')
import bench
print(bench.training_side_leg(torch.device('cuda')))
PY
