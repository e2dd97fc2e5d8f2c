cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
{
echo "== ncu steady-state dram traffic (single pass, caches untouched, rotating launches)"
timeout 600 ncu --cache-control none --clock-control none --launch-skip 60 --launch-count 16 -k regex:rw_main_kernel \
  --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --csv --log-file gpurun_out/r2_ncu_steady_traffic.csv \
  python bench.py --steps 40 --warmup 20 --no-cpu-baseline --no-extras --e2e-steps 1 --streams 1 > gpurun_out/r2_ncu_steady_traffic.out 2>&1
tail -3 gpurun_out/r2_ncu_steady_traffic.csv | cut -c1-300
echo "== ncu shared-memory bank conflicts (replayed passes)"
timeout 600 ncu --clock-control none --launch-skip 60 --launch-count 2 -k regex:rw_main_kernel \
  --metrics l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum,smsp__inst_executed.sum \
  --csv --log-file gpurun_out/r2_ncu_bank_conflicts.csv python bench.py --steps 10 --warmup 5 --no-cpu-baseline --no-extras --e2e-steps 1 --streams 1 > /dev/null 2>&1
tail -3 gpurun_out/r2_ncu_bank_conflicts.csv | cut -c1-400
echo "== launch list of the default bench command"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_bench_tiny4ag.csv python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/r2_launches.out 2>&1
python - <<'PY'
import csv, collections
rows = list(csv.reader(l for l in open("gpurun_out/r2_launches_bench_tiny4ag.csv") if l.startswith('"')))
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value")
tot = collections.Counter(); cnt = collections.Counter()
for r in rows[1:]:
    try: tot[r[ki][:60]] += float(r[vi].replace(",", "")); cnt[r[ki][:60]] += 1
    except Exception: pass
s = sum(tot.values())
for k, v in tot.most_common(8): print(f"{v/s*100:5.1f}%  n={cnt[k]:4d}  avg {v/cnt[k]/1e3:8.2f} us  {k}")
PY
for COOP in 0 1; do for TOOL in memcheck racecheck; do
echo "== compute-sanitizer $TOOL coop=$COOP"
RWARE_B200_COOP=$COOP timeout 900 compute-sanitizer --tool $TOOL --print-limit 20 python profiles/sanitize_workload.py > gpurun_out/r2_sanitizer_${TOOL}_coop${COOP}.log 2>&1
tail -4 gpurun_out/r2_sanitizer_${TOOL}_coop${COOP}.log
done; done
} > gpurun_out/job4.log 2>&1
