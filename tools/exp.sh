python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for w in cfg2 ddc16 cfg3 cfg5 cfg1; do python bench.py --workload $w --steps 10 --no-e2e --no-cpu 2>&1 | python tools/benchline.py | tail -1 | cut -c1-150; done
CMD="python bench.py --workload ddc16 --steps 1 --warmup 3 --no-e2e --no-cpu"
$CMD > gpurun_out/plain_tma_mix.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:dec_tma_kernel -s 3 -c 1 -f -o gpurun_out/prof_r1_dec_tma_mix_v1 $CMD > gpurun_out/ncu_tma_mix.log 2>&1
tail -2 gpurun_out/ncu_tma_mix.log
