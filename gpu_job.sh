cd $GRAFT_REPO_ROOT
timeout 200 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python bench.py --workload tiny-4ag-image --steps 100 --warmup 10 --no-cpu-baseline --e2e-steps 3 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('image', '%.3e'%d['value'], round(d['ms_per_step']*1000,2))"
