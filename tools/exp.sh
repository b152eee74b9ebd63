python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for w in cfg2 ddc16 cfg3; do python bench.py --workload $w --steps 10 --no-e2e --no-cpu 2>&1 | python tools/benchline.py | tail -1 | cut -c1-150; done
for r in 3 4 5; do echo raw=$r; SRCDSP_TMA_RAW=$r python bench.py --workload ddc16 --steps 10 --no-e2e --no-cpu 2>&1 | python tools/benchline.py | tail -1 | cut -c1-150; done
for r in 6 7; do echo cfg2 raw=$r; SRCDSP_TMA_RAW=$r python bench.py --workload cfg2 --steps 10 --no-e2e --no-cpu 2>&1 | python tools/benchline.py | tail -1 | cut -c1-150; done
