cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -3
python bench.py 2>&1 | tail -1 > gpurun_out/bench_r1.json; cat gpurun_out/bench_r1.json
