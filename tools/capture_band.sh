set -u
O=gpurun_out
cap() {
  local name=$1 rx=$2 skip=$3; shift 3
  local CMD="$*"
  if $CMD > $O/plain_$name.log 2>&1; then
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c 1 -f -o $O/prof_$name $CMD > $O/ncu_$name.log 2>&1
    timeout 300 ncu -i $O/prof_$name.ncu-rep --page details > $O/ncu_full_$name.txt 2>&1
    timeout 300 ncu -i $O/prof_$name.ncu-rep --page raw --csv --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__inst_executed.sum > $O/ncu_raw_$name.csv 2>&1
    timeout 300 ncu -i $O/prof_$name.ncu-rep --page source --csv --print-source sass > $O/ncu_source_$name.csv 2>/dev/null
    gzip -f $O/ncu_source_$name.csv
    rm -f $O/prof_$name.ncu-rep
  fi
}
B="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-ddc"
cap r2_dec_band_cfg2 dec_band_kernel 3 $B --workload cfg2
cap r2_dec_band_mix_ddc16 dec_band_kernel 3 $B --workload ddc16
