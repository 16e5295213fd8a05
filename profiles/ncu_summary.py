"""Key counters of every kernel in an .ncu-rep (run here, no GPU): python profiles/ncu_summary.py rep [out.csv]"""
import csv
import subprocess
import sys

raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], [r for r in rows[2:] if len(r) == len(rows[0])]
ix = {h: i for i, h in enumerate(hdr)}
WANT = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "smsp__inst_executed.sum",
        "sm__inst_executed.avg.per_cycle_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__waves_per_multiprocessor", "sm__cycles_active.avg", "sm__cycles_elapsed.max",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "sm__cycles_elapsed.avg.per_second"]
out = []
for r in data:
    rec = {w: r[ix[w]] for w in WANT if w in ix}
    st = [(h, float(r[i])) for h, i in ix.items() if h.startswith("smsp__average_warps_issue_stalled") and
          h.endswith("per_issue_active.ratio") and r[i] not in ("", "n/a")]
    rec["top_stalls"] = " ".join("%s:%.2f" % (h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), v)
                                 for h, v in sorted(st, key=lambda x: -x[1])[:7])
    out.append(rec)
    print("---- id", r[ix["ID"]])
    for k, v in rec.items():
        print("  %-75s %s" % (k, v))
if len(sys.argv) > 2:
    with open(sys.argv[2], "w", newline="") as f:
        w = csv.DictWriter(f, fieldnames=list(out[0].keys()))
        w.writeheader()
        w.writerows(out)
