#!/bin/bash
# Agent-split step (both warps of a pair, half of the agents each): GPU parity, then A/B against the previous build
# (robotic-warehouse_b200/lib/librware_b200_prev.so = HEAD before the change, same flags).
cd "$GRAFT_REPO_ROOT"
O=gpurun_out/pair_ab; mkdir -p $O
timeout 900 python -m pytest tests -x -q -m gpu > $O/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest.log
tail -3 $O/pytest.log
PREV=$PWD/robotic-warehouse_b200/lib/librware_b200_prev.so
for rep in 1 2; do
  timeout 300 python bench.py --steps 200 --warmup 20 > $O/new_$rep.json 2>> $O/err.log
  RWARE_B200_LIB_PATH=$PREV timeout 300 python bench.py --steps 200 --warmup 20 > $O/prev_$rep.json 2>> $O/err.log
done
python - <<'PY'
import json
for tag in ("new_1","prev_1","new_2","prev_2"):
    try:
        d=json.load(open(f"gpurun_out/pair_ab/{tag}.json"))
        print(tag, "us/step", round(d["ms_per_step"]*1e3,2), "frac", round(d["roofline"]["frac"],3), "single_set", round(d["sequential_rollout_of_one_world_set"]["ms_per_step"]*1e3,2),
              "e2e_ms", round(d["e2e"]["ms_per_step"],3), "others", [round(w["kernel_us"],1) for w in d.get("other_workloads",[])])
    except Exception as e:
        print(tag, "failed", e)
PY
