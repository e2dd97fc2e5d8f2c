cd $GRAFT_REPO_ROOT
RWARE_B200_HOST_TRACE=1 python profiles/hostexp/trace_step_host.py 2>&1 | tail -2
P='import sys,json; d=json.loads(sys.stdin.read()); e=d["e2e"]; print(sys.argv[1], "%.3e" % d["value"], "e2e %.3e" % e["value"], round(e["ms_per_step"],3), "ms; plain %.3e" % e["plain_float32_transport"]["value"], round(e["plain_float32_transport"]["ms_per_step"],3))'
for i in 1 2 3; do python bench.py --steps 50 --warmup 10 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "$P" run$i; done
for t in 6 12; do RWARE_B200_HOST_THREADS=$t python bench.py --steps 50 --warmup 10 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "$P" threads$t; done
for wl in medium-8ag-hard tiny-4ag-image; do python bench.py --workload $wl --steps 50 --warmup 10 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "$P" $wl; done
