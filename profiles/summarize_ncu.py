"""Summarise an .ncu-rep (brought back by gpurun) into the small JSON files committed under profiles/.

    python profiles/summarize_ncu.py gpurun_out/prof_step_r1f.ncu-rep profiles/r1_step_tiny4ag.json
"""
import csv
import json
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "launch__waves_per_multiprocessor", "sm__cycles_elapsed.max", "smsp__cycles_active.avg",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]


def main(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = {"kernel": r[hdr.index("Kernel Name")]}
        for w in WANT:
            if w in hdr:
                d[w] = {"value": r[hdr.index(w)], "unit": units[hdr.index(w)]}
        stalls = {h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""): float(r[i])
                  for i, h in enumerate(hdr)
                  if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio") and r[i]}
        d["stall_cycles_per_issued_instruction"] = dict(sorted(stalls.items(), key=lambda kv: -kv[1])[:8])
        res.append(d)
    with open(out, "w") as f:
        json.dump({"source": rep, "launches": res}, f, indent=1)
    print(out, len(res), "launch(es)")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
