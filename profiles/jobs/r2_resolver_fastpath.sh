#!/bin/bash
# Resolver fast path (all movers head for free, pairwise distinct cells): GPU suite + A/B against the previous build on one box.
cd "$GRAFT_REPO_ROOT"
O=gpurun_out/fastpath; mkdir -p $O
timeout 90 python -m pytest tests -x -q -m gpu > $O/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest.log
tail -2 $O/pytest.log
P='import sys,json; d=json.loads(sys.stdin.read()); print(sys.argv[1], "value %.3e"%d["value"], "us", round(d["ms_per_step"]*1000,2), "frac", round(d["roofline"]["frac"],3), "| single-set us", round(d["sequential_rollout_of_one_world_set"]["ms_per_step"]*1000,2), "| 2 streams", round(d["concurrent_streams"]["ms_per_step"]*1000,2))'
PREV=$PWD/robotic-warehouse_b200/lib/librware_b200_prev.so
for WL in tiny-4ag medium-8ag-hard; do
  timeout 45 python bench.py --workload $WL --steps 200 --warmup 20 --no-cpu-baseline --no-extras --e2e-steps 1 2>> $O/err.log | tail -1 | python -c "$P" new-$WL | tee -a $O/ab.txt
  RWARE_B200_LIB_PATH=$PREV timeout 45 python bench.py --workload $WL --steps 200 --warmup 20 --no-cpu-baseline --no-extras --e2e-steps 1 2>> $O/err.log | tail -1 | python -c "$P" prev-$WL | tee -a $O/ab.txt
done
