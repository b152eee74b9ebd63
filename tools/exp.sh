python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r1_cfg2.json 2> gpurun_out/bench_r1_cfg2.err; tail -c 400 gpurun_out/bench_r1_cfg2.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r1_reference.json 2>> gpurun_out/bench_r1_cfg2.err
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu"
$CMD > gpurun_out/plain_launches.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1_tma.csv $CMD > gpurun_out/ncu_launches.log 2>&1
python tools/benchline.py < gpurun_out/bench_r1_cfg2.json; python tools/benchline.py < gpurun_out/bench_r1_reference.json
