python tools/wbench.py
CMD="python bench.py --workload cfg4 --steps 1 --warmup 3 --no-e2e --no-cpu"
$CMD > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:up_fir_kernel -s 3 -c 1 -f -o gpurun_out/prof_r1_up_v1 $CMD > gpurun_out/ncu_up.log 2>&1
tail -1 gpurun_out/ncu_up.log
