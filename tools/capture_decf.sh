set -u
# float decimator evidence: bench lines, one ncu --set full capture of decf_fir_kernel (after the plain run exited 0), shapes
O=gpurun_out
python bench.py --workload cfg2f > $O/bench_cfg2f.json 2> $O/bench_cfg2f.err; tail -c 300 $O/bench_cfg2f.json; tail -3 $O/bench_cfg2f.err
python bench.py --workload cfg1f --no-cpu > $O/bench_cfg1f.json 2> $O/bench_cfg1f.err; tail -c 200 $O/bench_cfg1f.json
CMD="python tools/decfbench.py 16 255"
$CMD > $O/plain_decf.log 2>&1 && timeout 300 ncu --set full --clock-control none --import-source on -k regex:decf_fir_kernel -s 62 -c 1 -f -o $O/prof_r1_decf $CMD > $O/ncu_decf.log 2>&1
timeout 200 ncu -i $O/prof_r1_decf.ncu-rep --page details > $O/ncu_full_r1_decf.txt 2>&1
timeout 200 ncu -i $O/prof_r1_decf.ncu-rep --page raw --csv --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__inst_executed.sum,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio > $O/ncu_raw_r1_decf.csv 2>&1
tail -1 $O/ncu_raw_r1_decf.csv
for s in "16 255" "8 63" "4 1023" "2 32" "32 512"; do timeout 60 python tools/decfbench.py $s 2>&1 | tail -1; done | tee $O/decf_bench.log
