#!/bin/bash
# dwpw ablations (timing only): which phase owns the time?
python tools/bench_dwpw.py
for a in 1 2 3 4 5; do MAFB200_LIB=maf_yolo_b200/libmafb200_abl$a.so python tools/bench_dwpw.py; done
python tools/bench_dwpw.py
