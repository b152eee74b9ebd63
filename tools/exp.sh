python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for cfg in "4 3 0" "4 3 5" "4 2 3" "8 1 3"; do set -- $cfg; echo "ddc16 W=$1 groups=$2 stages=$3"; SRCDSP_TMA_W=$1 SRCDSP_TMA_GROUPS=$2 SRCDSP_TMA_STAGES=$3 python bench.py --workload ddc16 --steps 10 --no-e2e --no-cpu 2>&1 | python tools/benchline.py | tail -1 | cut -c1-120; done
python bench.py --workload cfg3 --steps 10 --no-e2e --no-cpu 2>&1 | python tools/benchline.py | tail -1 | cut -c1-150
