cd $GRAFT_REPO_ROOT
P='import sys,json; d=json.loads(sys.stdin.read()); print(sys.argv[1], d["value"], d["ms_per_step"], d["roofline"]["frac"], "obs_us", d["obs_gather"]["kernel_us"], d["obs_gather"]["frac_of_hbm_peak"])'
for parts in 2 4; do for wl in tiny-4ag tiny-2ag small-4ag medium-8ag-hard; do RWARE_B200_OBS_PARTS=$parts python bench.py --workload $wl --steps 200 --warmup 20 --no-cpu-baseline --e2e-steps 2 2>&1 | tail -1 | python -c "$P" parts$parts-$wl; done; done
RWARE_B200_OBS_PARTS=4 timeout 900 python -m pytest tests/test_hostsim_golden.py tests/test_known_answers.py -m gpu -q 2>&1 | tail -1
RWARE_B200_OBS_PARTS=4 python profiles/phase_timeline.py 2>&1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
for m in d: print(m, {k[5]+k[9:12]: (round(v,1) if isinstance(v,float) else v) for k,v in d[m]['primary'].items() if 'mean' in k}, 'span', round(d[m]['kernel_span_us'],1))
"
