set -x
cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -4
python bench.py --steps 100 --warmup 10 --no-cpu-baseline --e2e-steps 3 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('STATIC', d['value'], d['ms_per_step'], d['value_warm_l2'], d['roofline']['frac'])"
RWARE_B200_FORCE_DYNAMIC=1 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --e2e-steps 3 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('DYNAMIC', d['value'], d['ms_per_step'], d['value_warm_l2'])"
RWARE_B200_OBS_NBUF=2 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --e2e-steps 3 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('NBUF2', d['value'], d['ms_per_step'], d['value_warm_l2'])"
for w in 4 8 12; do RWARE_B200_WARPS_PER_CTA=$w python bench.py --steps 100 --warmup 10 --no-cpu-baseline --e2e-steps 3 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('WPC$w', d['value'], d['ms_per_step'], d['value_warm_l2'])"; done
ncu --set full --clock-control none --import-source on -k regex:rw_main_kernel -s 6 -c 1 -o gpurun_out/prof_step_r1b python bench.py --steps 5 --warmup 5 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_full.log 2>&1
