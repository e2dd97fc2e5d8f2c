set -x
cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -4
P='import sys,json; d=json.loads(sys.stdin.read()); print(sys.argv[1], d["value"], d["ms_per_step"], d["value_warm_l2"], d["roofline"]["frac"], "obs_us", d["obs_gather"]["kernel_us"], d["obs_gather"]["frac_of_hbm_peak"])'
python bench.py --steps 100 --warmup 10 --no-cpu-baseline --e2e-steps 3 2>&1 | tail -1 | python -c "$P" FULLSTATIC
for wl in tiny-2ag small-4ag medium-8ag-hard large-16ag-r2; do python bench.py --workload $wl --steps 50 --warmup 10 --no-cpu-baseline --e2e-steps 2 2>&1 | tail -1 | python -c "$P" $wl; done
ncu --set full --clock-control none --import-source on -k regex:rw_main_kernel -s 6 -c 1 -o gpurun_out/prof_step_r1d python bench.py --steps 5 --warmup 5 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_full.log 2>&1
