set -u
O=gpurun_out
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke_final.log 2>&1; tail -1 $O/smoke_final.log
python bench.py > $O/bench_default.json 2> $O/bench_default.err; tail -c 300 $O/bench_default.json
python bench.py --impl reference > $O/bench_reference.json 2> $O/bench_reference.err; tail -c 200 $O/bench_reference.json
CMD="python tools/upbench.py 8 96"
$CMD > $O/plain_up_tc.log 2>&1 && timeout 300 ncu --set full --clock-control none --import-source on -k regex:up_tc_kernel -s 3 -c 1 -f -o $O/prof_r1_up_tc $CMD > $O/ncu_up_tc.log 2>&1
timeout 200 ncu -i $O/prof_r1_up_tc.ncu-rep --page details > $O/ncu_full_r1_up_tc.txt 2>&1
timeout 200 ncu -i $O/prof_r1_up_tc.ncu-rep --page raw --csv --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__inst_executed.sum > $O/ncu_raw_r1_up_tc.csv 2>&1
tail -1 $O/ncu_raw_r1_up_tc.csv
