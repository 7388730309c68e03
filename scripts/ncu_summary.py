"""Summarise ncu output for profiles/: launch-list shares (csv from --metrics gpu__time_duration.sum) and key metrics
of a --set full capture (raw page csv piped in)."""
import csv, sys, collections, subprocess

def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = [i for i, r in enumerate(rows) if r[0] == "ID"][0]
    H = rows[hdr]; data = rows[hdr + 1:]
    ik, iv = H.index("Kernel Name"), H.index("Metric Value")
    tot = collections.defaultdict(float); cnt = collections.Counter()
    for r in data:
        name = r[ik].split("(")[0]; v = float(r[iv].replace(",", ""))
        tot[name] += v; cnt[name] += 1
    s = sum(tot.values())
    out = ["kernel,launches,total_us,share_pct"]
    for k, v in sorted(tot.items(), key=lambda x: -x[1]):
        out.append(f"{k},{cnt[k]},{v/1e3:.1f},{v/s*100:.1f}")
    return "\n".join(out)

KEEP = ["Kernel Name", "launch__grid_size", "launch__registers_per_thread", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio"]

def full(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    H = rows[0]
    idx = [H.index(k) for k in KEEP if k in H]
    out = [",".join(H[i] for i in idx), ",".join(rows[1][i] for i in idx)]
    for r in rows[2:]:
        out.append(",".join('"' + r[i].split("(")[0][:40] + '"' if H[i] == "Kernel Name" else r[i] for i in idx))
    return "\n".join(out)

if __name__ == "__main__":
    if sys.argv[1] == "launches":
        print(launches(sys.argv[2]))
    else:
        print(full(sys.argv[2]))
