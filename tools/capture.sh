#!/bin/bash
# Round captures for profiles/: ncu --set full of the dominant kernel of each workload (only after the same command
# has exited 0 without ncu), the launch list of the default bench command, and the bench lines themselves.
# Run under gpurun from the repo root (tools/capture.sh [round-tag]); everything lands in gpurun_out/.
set -u
O=gpurun_out
R=${1:-r2}
cap() {  # cap <name> <kernel regex> <skip> <command...>
  local name=$1 rx=$2 skip=$3; shift 3
  local CMD="$*"
  if $CMD > $O/plain_$name.log 2>&1; then
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c 1 -f -o $O/prof_$name $CMD > $O/ncu_$name.log 2>&1
    timeout 300 ncu -i $O/prof_$name.ncu-rep --page details > $O/ncu_full_$name.txt 2>&1
    timeout 300 ncu -i $O/prof_$name.ncu-rep --page raw --csv --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__inst_executed.sum > $O/ncu_raw_$name.csv 2>&1
    # per-instruction counters (executed counts, shared-memory wavefronts, stall samples) instead of the 14 MB report:
    # gpurun brings back at most 64 MiB
    timeout 300 ncu -i $O/prof_$name.ncu-rep --page source --csv --print-source sass > $O/ncu_source_$name.csv 2>/dev/null
    gzip -f $O/ncu_source_$name.csv
    rm -f $O/prof_$name.ncu-rep
  else
    echo "plain run of $name failed" >&2
  fi
}
B="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-ddc"
cap ${R}_dec_band_cfg2 dec_band_kernel 3 $B --workload cfg2
cap ${R}_dec_band_mix_ddc16 dec_band_kernel 3 $B --workload ddc16
cap ${R}_dec_tma_p2_cfg5 dec_tma_kernel 1 $B --workload cfg5
cap ${R}_dec_band_cfg3_stage1 dec_band_kernel 3 $B --workload cfg3
cap ${R}_dec_tma_cfg3_stage2 dec_tma_kernel 3 $B --workload cfg3
cap ${R}_up_fir4_cfg4 up_fir4_kernel 3 $B --workload cfg4
CMD="python tools/decfbench.py 16 255"
$CMD > $O/plain_${R}_decf.log 2>&1 && timeout 300 ncu --set full --clock-control none --import-source on -k regex:decf_fir_kernel -s 62 -c 1 -f -o $O/prof_${R}_decf $CMD > $O/ncu_${R}_decf.log 2>&1
timeout 200 ncu -i $O/prof_${R}_decf.ncu-rep --page details > $O/ncu_full_${R}_decf.txt 2>&1
timeout 200 ncu -i $O/prof_${R}_decf.ncu-rep --page raw --csv --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__inst_executed.sum,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active > $O/ncu_raw_${R}_decf.csv 2>&1
# launch list of the default bench command (short)
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu"
$CMD > $O/plain_launches.log 2>&1 && timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_${R}.csv $CMD > $O/ncu_launches.log 2>&1
rm -f $O/prof_${R}_decf.ncu-rep
