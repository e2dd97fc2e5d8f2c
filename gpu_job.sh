cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -2
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py 2>&1 | tail -1 > gpurun_out/bench_r1.json; cut -c1-260 gpurun_out/bench_r1.json
python bench.py --impl reference --steps 1 --warmup 0 2>&1 | tail -1 > gpurun_out/bench_ref_r1.json; cut -c1-200 gpurun_out/bench_ref_r1.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 10 --warmup 5 --no-cpu-baseline --e2e-steps 2 > gpurun_out/bench_under_ncu.log 2>&1
grep -c rw_main_kernel gpurun_out/launches_r1.csv
ncu --set full --clock-control none --import-source on -k regex:rw_main_kernel -s 4 -c 1 -o gpurun_out/prof_step_r1_final python profiles/prof_step.py > gpurun_out/ncu_final.log 2>&1
python profiles/phase_timeline.py > gpurun_out/phase_timeline_tiny4ag.json 2>&1; tail -3 gpurun_out/phase_timeline_tiny4ag.json
