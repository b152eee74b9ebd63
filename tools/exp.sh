python -m pytest tests -m gpu -x -q 2>&1 | tail -5
B="python bench.py --workload cfg2 --steps 10 --no-e2e --no-cpu"
for c in 4 6 8 12 16; do echo "conv=$c"; SRCDSP_TMA_CONV=$c $B 2>&1 | python tools/benchline.py; done
for r in 3 4 5 6; do echo "raw=$r"; SRCDSP_TMA_RAW=$r $B 2>&1 | python tools/benchline.py; done
for d in 1 4 5 2 3 7; do echo "debug=$d"; SRCDSP_TC_DEBUG=$d $B 2>&1 | python tools/benchline.py; done
echo counters; SRCDSP_TC_DEBUG=32 python bench.py --workload cfg2 --steps 3 --no-e2e --no-cpu 2>&1 | tail -3
