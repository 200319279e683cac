mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:nms -s 2 -c 2 --csv --log-file gpurun_out/nms25.csv python tools/one_forward.py > gpurun_out/ncu25.log 2>&1
grep -E "nms_" gpurun_out/nms25.csv | awk -F'","' '{print $5, $(NF)}'
timeout 600 ncu --set full --clock-control none --import-source on -k regex:nms_select -s 1 -c 1 -o gpurun_out/prof25_nms_select -f python tools/one_forward.py > gpurun_out/ncu25b.log 2>&1
