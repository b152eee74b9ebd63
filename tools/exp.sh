python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for w in cfg2 ddc16 cfg3; do python bench.py --workload $w --steps 10 --no-e2e --no-cpu 2>&1 | python tools/benchline.py | tail -1 | cut -c1-150; done
for cfg in "4 2 3" "4 4 5" "8 2 3" "8 2 4"; do set -- $cfg; echo "ddc16 W=$1 groups=$2 stages=$3"; SRCDSP_TMA_W=$1 SRCDSP_TMA_GROUPS=$2 SRCDSP_TMA_STAGES=$3 python bench.py --workload ddc16 --steps 10 --no-e2e --no-cpu 2>&1 | python tools/benchline.py | tail -1 | cut -c1-120; done
python bench.py --workload cfg2 --kernel 3 --steps 10 --no-e2e --no-cpu 2>&1 | python tools/benchline.py | tail -1 | cut -c1-150
