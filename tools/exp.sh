python -m pytest tests -m gpu -x -q 2>&1 | tail -12
python bench.py --workload fifo --steps 2 --warmup 3 2>&1 | python tools/benchline.py | tail -2 | cut -c1-300
