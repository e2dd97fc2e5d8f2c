cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
{
echo "== gpu tests"
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
echo "== kernel A/B: lib_old (round-2 start) vs lib (current)"
P='import sys,json; d=json.loads(sys.stdin.read()); print(sys.argv[1], "value %.3e"%d["value"], "us", round(d["ms_per_step"]*1000,2), "frac", round(d["roofline"]["frac"],3), "| obs us", round(d["obs_gather"]["kernel_us"],2), "| single-set us", round(d["sequential_rollout_of_one_world_set"]["ms_per_step"]*1000,2) if "sequential_rollout_of_one_world_set" in d else "")'
for ROUND in 1 2; do for WL in tiny-4ag medium-8ag-hard large-16ag-r2 small-4ag; do for LIB in lib_old lib; do
RWARE_B200_LIB_PATH=$PWD/robotic-warehouse_b200/$LIB/librware_b200.so python bench.py --workload $WL --steps 200 --warmup 20 --no-cpu-baseline --no-extras --e2e-steps 1 --streams 1 2>/dev/null | tail -1 | python -c "$P" $WL-$LIB-r$ROUND
done; done; done
echo "== sweep"
timeout 300 python profiles/hostexp/e2e_sweep.py rware-tiny-4ag-v1 65536 quick
echo "== bench with host trace"
RWARE_B200_HOST_TRACE=1 timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/job8_bench.out 2> gpurun_out/job8_bench.err
tail -1 gpurun_out/job8_bench.out
grep "rw_step_host" gpurun_out/job8_bench.err | tail -30
echo "== racecheck coop=0"
timeout 900 compute-sanitizer --tool racecheck --print-limit 300 python profiles/sanitize_workload.py > gpurun_out/r2_sanitizer_racecheck_coop0.log 2>&1
tail -2 gpurun_out/r2_sanitizer_racecheck_coop0.log
} > gpurun_out/job8.log 2>&1
