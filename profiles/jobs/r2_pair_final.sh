#!/bin/bash
# Agent-split step incl. 16 agents: GPU suite, compute-sanitizer (memcheck + racecheck) on profiles/sanitize_workload.py,
# then the default bench of this build and of the previous build (librware_b200_prev.so: one-warp step) on the same box.
cd "$GRAFT_REPO_ROOT"
O=gpurun_out/pair_final; mkdir -p $O
timeout 600 python -m pytest tests -x -q -m gpu > $O/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest.log
tail -3 $O/pytest.log
timeout 300 python bench.py --steps 200 --warmup 20 > $O/new.json 2>> $O/err.log; echo "bench rc=$?"
RWARE_B200_LIB_PATH=$PWD/robotic-warehouse_b200/lib/librware_b200_prev.so timeout 300 python bench.py --steps 200 --warmup 20 --no-extras > $O/prev_noextras.json 2>> $O/err.log
RWARE_B200_LIB_PATH=$PWD/robotic-warehouse_b200/lib/librware_b200_prev.so timeout 300 python bench.py --steps 200 --warmup 20 > $O/prev.json 2>> $O/err.log
python - <<'PY'
import json
for tag in ("new","prev"):
    try:
        d=json.load(open(f"gpurun_out/pair_final/{tag}.json"))
        print(tag, "us/step", round(d["ms_per_step"]*1e3,2), "frac", round(d["roofline"]["frac"],3), "single_set", round(d["sequential_rollout_of_one_world_set"]["ms_per_step"]*1e3,2),
              "e2e_ms", round(d["e2e"]["ms_per_step"],3), "others", [round(w["kernel_us"],1) for w in d.get("other_workloads",[])])
    except Exception as e:
        print(tag, "failed", e)
PY
for TOOL in memcheck racecheck; do
  SAN_ENVS=200 timeout 200 compute-sanitizer --tool $TOOL --print-limit 20 python profiles/sanitize_workload.py > $O/sanitizer_${TOOL}.log 2>&1; echo "$TOOL rc=$?"
  tail -3 $O/sanitizer_${TOOL}.log
done
