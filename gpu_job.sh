cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_views.py -m gpu -x -q 2>&1 | tail -3
timeout 300 python profiles/prof_views.py > gpurun_out/views.json 2> gpurun_out/views.err; tail -3 gpurun_out/views.err; python -c "
import json; d=json.load(open('gpurun_out/views.json'))
for c in d['cases']: print(c['kernel'], c['env'], c.get('layers'), c.get('pad_to_shape'), c['us_per_launch'], 'us', c['gbs'], 'GB/s', c['frac_of_peak'])
"
