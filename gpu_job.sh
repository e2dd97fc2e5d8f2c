cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_obs_types.py tests/test_hostsim_golden.py -m gpu -x -q 2>&1 | tail -2
P='import sys,json; d=json.loads(sys.stdin.read()); print(sys.argv[1], "%.3e" % d["value"], round(d["ms_per_step"]*1000,2), "us frac", round(d["roofline"]["frac"],3), "obs_us", round(d["obs_gather"]["kernel_us"],2), "multi", d.get("concurrent_streams"))'
for wl in tiny-4ag-image tiny-4ag-imagedict; do python bench.py --workload $wl --steps 200 --warmup 20 --no-cpu-baseline --e2e-steps 2 2>&1 | tail -1 | python -c "$P" $wl; done
for wpc in 14 7 5; do for st in 2 3; do RWARE_B200_WARPS_PER_CTA=$wpc python bench.py --streams $st --steps 200 --warmup 20 --no-cpu-baseline --e2e-steps 2 2>&1 | tail -1 | python -c "$P" wpc$wpc-streams$st; done; done
RWARE_B200_WARPS_PER_CTA=7 python bench.py --workload medium-8ag-hard --streams 2 --steps 200 --warmup 20 --no-cpu-baseline --e2e-steps 2 2>&1 | tail -1 | python -c "$P" medium-wpc7-streams2
RWARE_B200_WARPS_PER_CTA=2 python bench.py --workload medium-8ag-hard --streams 2 --steps 200 --warmup 20 --no-cpu-baseline --e2e-steps 2 2>&1 | tail -1 | python -c "$P" medium-wpc2-streams2
