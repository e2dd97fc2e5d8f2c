cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
P='import sys,json; d=json.loads(sys.stdin.read()); print(sys.argv[1], "%.3e" % d["value"], round(d["ms_per_step"]*1000,2), "us frac", d["roofline"]["frac"], "obs_us", d["obs_gather"]["kernel_us"], d["obs_gather"]["frac_of_hbm_peak"])'
for wl in tiny-4ag tiny-2ag small-4ag medium-8ag-hard large-16ag-r2; do python bench.py --workload $wl --steps 200 --warmup 20 --no-cpu-baseline --e2e-steps 2 2>&1 | tail -1 | python -c "$P" $wl; done
for wl in medium-8ag-hard large-16ag-r2; do RWARE_B200_SMALL_CTA=0 python bench.py --workload $wl --steps 200 --warmup 20 --no-cpu-baseline --e2e-steps 2 2>&1 | tail -1 | python -c "$P" nosmall-$wl; done
python profiles/phase_timeline.py 2>&1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
for m in d: print(m, {k[5]+k[9:12]: (round(v,1) if isinstance(v,float) else v) for k,v in d[m]['primary'].items() if 'mean' in k}, 'span', round(d[m]['kernel_span_us'],1))
"
