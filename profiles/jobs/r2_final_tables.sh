#!/bin/bash
# Final state of round 2: GPU suite, smoke, the default bench line (both protocols: 200 launches and the driver's 20), the per-workload table.
cd "$GRAFT_REPO_ROOT"
O=gpurun_out/final2; mkdir -p $O
timeout 600 python -m pytest tests -x -q -m gpu > $O/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest.log
tail -2 $O/pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 300 python bench.py > $O/bench.json 2>> $O/err.log; echo "bench rc=$?"
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras > $O/bench_20.json 2>> $O/err.log
P='import sys,json; d=json.loads(sys.stdin.read()); print(sys.argv[1], "value %.3e"%d["value"], "us", round(d["ms_per_step"]*1000,2), "frac", round(d["roofline"]["frac"],3), "| obs us", round(d["obs_gather"]["kernel_us"],2), "obs frac", round(d["obs_gather"]["frac_of_hbm_peak"],3), "| single-set us", round(d["sequential_rollout_of_one_world_set"]["ms_per_step"]*1000,2), "| 2 streams", round(d["concurrent_streams"]["ms_per_step"]*1000,2))'
echo "# python bench.py --workload W --steps 200 --warmup 20 --no-cpu-baseline --no-extras --e2e-steps 1  (profiles/jobs/r2_final_tables.sh, one box; agent-split step)" > $O/workloads.txt
for WL in tiny-4ag tiny-2ag small-4ag medium-8ag-hard large-16ag-r2 tiny-4ag-image; do
  timeout 200 python bench.py --workload $WL --steps 200 --warmup 20 --no-cpu-baseline --no-extras --e2e-steps 1 2>> $O/err.log | tail -1 | python -c "$P" $WL >> $O/workloads.txt
done
cat $O/workloads.txt
python - <<'PY'
import json
for tag in ("bench","bench_20"):
    d=json.load(open(f"gpurun_out/final2/{tag}.json"))
    print(tag, "us/step", round(d["ms_per_step"]*1e3,2), "value %.3e"%d["value"], "frac", round(d["roofline"]["frac"],3), "e2e_ms", round(d["e2e"]["ms_per_step"],3), "e2e value %.3e"%d["e2e"]["value"])
PY
