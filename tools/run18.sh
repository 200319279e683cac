mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum,launch__registers_per_thread,launch__occupancy_limit_shared_mem,launch__occupancy_limit_registers,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:dwconv -s 16 -c 16 --csv --log-file gpurun_out/dw18.csv python tools/one_forward.py > gpurun_out/ncu18.log 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/dw18.csv")) if len(r)>10]
hdr=rows[0]; 
from collections import defaultdict
d=defaultdict(dict)
for r in rows[1:]:
    d[r[hdr.index("ID")]]["name"]=r[hdr.index("Kernel Name")][:40]; d[r[hdr.index("ID")]]["grid"]=r[hdr.index("Grid Size")]
    d[r[hdr.index("ID")]][r[hdr.index("Metric Name")]]=r[hdr.index("Metric Value")]
for k,v in d.items(): print(k, v)
PY
