#!/bin/bash
# ncu --set full of the float decimator alone (tools/decfbench.py 16 255); run under gpurun: tools/capture_decf.sh [tag]
set -u
O=gpurun_out
R=${1:-r2b}
CMD="python tools/decfbench.py 16 255"
$CMD > $O/plain_${R}_decf.log 2>&1 && timeout 300 ncu --set full --clock-control none --import-source on -k regex:decf_quad_kernel -s 20 -c 1 -f -o $O/prof_${R}_decf $CMD > $O/ncu_${R}_decf.log 2>&1
timeout 200 ncu -i $O/prof_${R}_decf.ncu-rep --page details > $O/ncu_full_${R}_decf.txt 2>&1
timeout 200 ncu -i $O/prof_${R}_decf.ncu-rep --page raw --csv --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__inst_executed.sum,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active > $O/ncu_raw_${R}_decf.csv 2>&1
timeout 300 ncu -i $O/prof_${R}_decf.ncu-rep --page source --csv --print-source sass > $O/ncu_source_${R}_decf.csv 2>/dev/null
gzip -f $O/ncu_source_${R}_decf.csv
rm -f $O/prof_${R}_decf.ncu-rep
tail -2 $O/plain_${R}_decf.log
