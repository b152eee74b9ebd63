set -u
# round-end check on one B200: GPU tests, smoke, default bench line + reference arm, float workload, one ncu capture
O=gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > $O/pytest_gpu_final.log 2>&1; tail -3 $O/pytest_gpu_final.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke_final.log 2>&1; tail -1 $O/smoke_final.log
python bench.py > $O/bench_default.json 2> $O/bench_default.err; tail -c 300 $O/bench_default.json
python bench.py --impl reference > $O/bench_reference.json 2> $O/bench_reference.err; tail -c 200 $O/bench_reference.json
python bench.py --workload cfg2f > $O/bench_cfg2f.json 2> $O/bench_cfg2f.err; tail -c 300 $O/bench_cfg2f.json
python bench.py --workload cfg1f --no-cpu > $O/bench_cfg1f.json 2> $O/bench_cfg1f.err; tail -c 200 $O/bench_cfg1f.json
CMD="python tools/decfbench.py 16 255"
$CMD > $O/plain_decf.log 2>&1 && timeout 300 ncu --set full --clock-control none --import-source on -k regex:decf_fir_kernel -s 36 -c 1 -f -o $O/prof_r1_decf $CMD > $O/ncu_decf.log 2>&1
timeout 200 ncu -i $O/prof_r1_decf.ncu-rep --page details > $O/ncu_full_r1_decf.txt 2>&1
timeout 200 ncu -i $O/prof_r1_decf.ncu-rep --page raw --csv --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__inst_executed.sum > $O/ncu_raw_r1_decf.csv 2>&1
tail -1 $O/ncu_raw_r1_decf.csv
cat $O/plain_decf.log
