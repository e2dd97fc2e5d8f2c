cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -1
P='import sys,json; d=json.loads(sys.stdin.read()); e=d["e2e"]; print(sys.argv[1], "%.3e" % d["value"], round(d["ms_per_step"]*1000,2), "us frac", round(d["roofline"]["frac"],3), "e2e %.3e" % e["value"], "streams2 %.3e" % d["concurrent_streams"]["value"])'
python bench.py --workload medium-8ag-hard --p-forward 0.68 --steps 200 --warmup 20 --no-cpu-baseline --e2e-steps 4 2>/dev/null | tail -1 | python -c "$P" medium-pf0.68
python bench.py --workload tiny-4ag --p-forward 0.68 --steps 200 --warmup 20 --no-cpu-baseline --e2e-steps 4 2>/dev/null | tail -1 | python -c "$P" tiny4-pf0.68
python bench.py --workload large-16ag-r2 --p-forward 0.68 --steps 200 --warmup 20 --no-cpu-baseline --e2e-steps 4 2>/dev/null | tail -1 | python -c "$P" large-pf0.68
