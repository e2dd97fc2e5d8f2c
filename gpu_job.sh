cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -2
python profiles/phase_timeline.py 2>&1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
for m in d: print(m, {k[5]+k[9:12]: (round(v,1) if isinstance(v,float) else v) for k,v in d[m]['primary'].items() if 'mean' in k}, 'helper', {k[5]: round(v,1) for k,v in d[m]['helper'].items() if v is not None and 'mean' in k}, 'span', round(d[m]['kernel_span_us'],1))
"
P='import sys,json; d=json.loads(sys.stdin.read()); print(sys.argv[1], d["value"], d["ms_per_step"], d["value_single_set_l2_resident"], d["roofline"]["frac"], "obs_us", d["obs_gather"]["kernel_us"], d["obs_gather"]["frac_of_hbm_peak"])'
python bench.py --steps 200 --warmup 20 --no-cpu-baseline --e2e-steps 3 2>&1 | tail -1 | python -c "$P" PDL1
RWARE_B200_PDL=0 python bench.py --steps 200 --warmup 20 --no-cpu-baseline --e2e-steps 3 2>&1 | tail -1 | python -c "$P" PDL0
for wl in tiny-2ag small-4ag medium-8ag-hard large-16ag-r2; do python bench.py --workload $wl --steps 100 --warmup 10 --no-cpu-baseline --e2e-steps 2 2>&1 | tail -1 | python -c "$P" $wl; done
