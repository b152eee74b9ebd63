python -m pytest tests -m gpu -x -q 2>&1 | tail -3
B="python bench.py --workload cfg2 --steps 10 --no-e2e --no-cpu"
$B 2>&1 | python tools/benchline.py
for d in 10 14 1 2 4; do echo "debug=$d"; SRCDSP_TC_DEBUG=$d $B 2>&1 | python tools/benchline.py; done
for cg in "8 2" "12 2" "12 3" "16 1"; do set -- $cg; echo "conv=$1 groups=$2"; SRCDSP_TMA_CONV=$1 SRCDSP_TMA_GROUPS=$2 $B 2>&1 | python tools/benchline.py; done
python bench.py --workload cfg2 --kernel 3 --steps 10 --no-e2e --no-cpu 2>&1 | python tools/benchline.py
python bench.py --workload ddc16 --steps 10 --no-e2e --no-cpu 2>&1 | python tools/benchline.py
SRCDSP_TC_DEBUG=32 python bench.py --workload cfg2 --steps 3 --no-e2e --no-cpu 2>&1 | tail -2 | head -1
