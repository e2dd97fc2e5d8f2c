set -x
cd $GRAFT_REPO_ROOT
nproc
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -3
python bench.py --impl reference --steps 1 --warmup 0 2>&1 | tail -1 > gpurun_out/bench_ref_r1.json; cat gpurun_out/bench_ref_r1.json | cut -c1-400
python bench.py 2>&1 | tail -1 > gpurun_out/bench_r1.json; cat gpurun_out/bench_r1.json
ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 60 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 10 --warmup 5 --no-cpu-baseline --e2e-steps 2 > gpurun_out/bench_under_ncu.log 2>&1
tail -5 gpurun_out/launches_r1.csv
