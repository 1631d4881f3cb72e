"""Print the judge-relevant metrics of an ncu --set full report (one kernel launch) as plain text."""
import csv
import subprocess
import sys

rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[0]
want = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__t_bytes.sum", "sm__cycles_elapsed.max",
        "smsp__inst_executed.sum"]
for r in rows[2:]:
    print("=" * 100)
    d = dict(zip(hdr, r))
    for k in want:
        if k in d:
            print("%-75s %s %s" % (k, d[k], rows[1][hdr.index(k)] if len(rows) > 1 else ""))
    stalls = sorted(((float(v.replace(",", "")), k) for k, v in d.items() if k.startswith("smsp__pcsamp_warps_issue_stalled_") and not k.endswith("not_issued") and v not in ("", "n/a")), reverse=True)
    tot = sum(v for v, _ in stalls) or 1.0
    print("warp stall samples: " + ", ".join("%s %.1f%%" % (k.replace("smsp__pcsamp_warps_issue_stalled_", ""), 100 * v / tot) for v, k in stalls[:8]))
