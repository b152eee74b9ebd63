set -u
# round-end check on one B200: GPU tests, smoke, default bench line + reference arm, float workloads
O=gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > $O/pytest_gpu_final.log 2>&1; tail -3 $O/pytest_gpu_final.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke_final.log 2>&1; tail -1 $O/smoke_final.log
python bench.py > $O/bench_default.json 2> $O/bench_default.err; tail -c 300 $O/bench_default.json
python bench.py --impl reference > $O/bench_reference.json 2> $O/bench_reference.err; tail -c 200 $O/bench_reference.json
python bench.py --workload cfg2f > $O/bench_cfg2f.json 2> $O/bench_cfg2f.err; tail -c 300 $O/bench_cfg2f.json
python bench.py --workload cfg1f --no-cpu > $O/bench_cfg1f.json 2> $O/bench_cfg1f.err; tail -c 200 $O/bench_cfg1f.json
