#!/bin/bash
# Round captures for profiles/: ncu --set full of the dominant kernel of each workload (only after the same command
# has exited 0 without ncu), the launch list of the default bench command, and the bench lines themselves.
# Run under gpurun from the repo root; everything lands in gpurun_out/.
set -u
O=gpurun_out
cap() {  # cap <name> <kernel regex> <skip> <bench args...>
  local name=$1 rx=$2 skip=$3; shift 3
  local CMD="python bench.py $* --steps 1 --warmup 3 --no-e2e --no-cpu --no-ddc"
  if $CMD > $O/plain_$name.log 2>&1; then
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c 1 -f -o $O/prof_$name $CMD > $O/ncu_$name.log 2>&1
    timeout 300 ncu -i $O/prof_$name.ncu-rep --page details > $O/ncu_full_$name.txt 2>&1
    timeout 300 ncu -i $O/prof_$name.ncu-rep --page raw --csv --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__inst_executed.sum > $O/ncu_raw_$name.csv 2>&1
  else
    echo "plain run of $name failed" >&2
  fi
}
cap r1_dec_tma_cfg2 dec_tma_kernel 3 --workload cfg2
cap r1_dec_tma_mix_ddc16 dec_tma_kernel 3 --workload ddc16
cap r1_dec_tma_p2_cfg5 dec_tma_kernel 3 --workload cfg5
cap r1_up_fir4_cfg4 up_fir4_kernel 3 --workload cfg4
cap r1_corr_blocked corr_scan_blocked_kernel 3 --workload corr
# launch list of the default bench command (short)
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu"
$CMD > $O/plain_launches.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_r1_final.csv $CMD > $O/ncu_launches.log 2>&1
# bench lines
python bench.py > $O/bench_default.json 2> $O/bench_default.err
python bench.py --impl reference > $O/bench_reference.json 2> $O/bench_reference.err
: > $O/bench_other.jsonl
for w in ddc16 ddc8 cfg3 cfg5 cfg4 mix corr cfg1; do python bench.py --workload $w --steps 10 --warmup 3 --no-e2e --no-cpu --no-ddc 2>/dev/null | tail -1 >> $O/bench_other.jsonl; done
