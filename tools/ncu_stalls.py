"""Key metrics + stall-reason shares of every launch in an ncu report -> CSV (development aid; used for the captures in
profiles/ that are not part of tools/gpu_round.sh).  usage: python tools/ncu_stalls.py <report.ncu-rep> <out.csv>"""
import csv, io, re, subprocess, sys

rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
KEYS = [("gpu__time_duration.sum", "duration"), ("dram__bytes_read.sum", "dram_read"), ("dram__bytes_write.sum", "dram_write"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct_of_peak"),
        ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex_pct_of_peak"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_pct_of_peak"),
        ("l1tex__t_sector_hit_rate.pct", "l1_hit_pct"), ("lts__t_sector_hit_rate.pct", "l2_hit_pct"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"),
        ("smsp__thread_inst_executed_per_inst_executed.ratio", "active_lanes_per_inst"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved_occupancy_pct"),
        ("launch__registers_per_thread", "registers"), ("smsp__inst_executed.sum", "warp_insts")]
STALL = "smsp__pcsamp_warps_issue_stalled_"
stall_cols = [h for h in hdr if h.startswith(STALL) and "not_issued" not in h]
with open(out, "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["kernel"] + [f"{s} [{units[idx[m]]}]" for m, s in KEYS if m in idx] + ["stall_" + h[len(STALL):] + "_pct" for h in stall_cols])
    for d in data:
        tot = sum(float(d[idx[h]].replace(",", "")) for h in stall_cols) or 1.0
        w.writerow([re.sub(r"\(.*", "", d[idx["Kernel Name"]])] + [d[idx[m]] for m, s in KEYS if m in idx] +
                   [round(100 * float(d[idx[h]].replace(",", "")) / tot, 1) for h in stall_cols])
print("wrote", out, len(data), "launches")
