"""Turn the ncu artefacts of a gpurun call (gpurun_out/prof_*_raw.csv = `ncu --page raw --csv` of each capture, written on
the GPU box by tools/gpu_round.sh; launches_bench.csv) into the committed summaries under profiles/.
Usage: python tools/ncu_summary.py r02"""
import collections, csv, io, json, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"

METRICS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram_read"),
    ("dram__bytes_write.sum", "dram_write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct_of_peak"),
    ("l1tex__m_xbar2l1tex_read_bytes.sum", "l2_bytes"),   # bytes the SMs read from L2 (L2 -> L1 crossbar)
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_pct_of_peak"),
    ("lts__t_sector_hit_rate.pct", "l2_hit_pct"),
    ("l1tex__t_sector_hit_rate.pct", "l1_hit_pct"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1_pct_of_peak"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct_of_peak"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active_threads_per_inst"),
    ("smsp__inst_executed.sum", "warp_insts"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved_occupancy_pct"),
    ("launch__registers_per_thread", "registers"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("smsp__warps_eligible.avg.per_cycle_active", "eligible_warps_per_cycle"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64_pipe_pct"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall_long_scoreboard"),
]


def raw(path):
    if path.endswith(".ncu-rep"):
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    else:
        out = open(path).read()
    rows = [r for r in csv.reader(io.StringIO(out)) if r and not r[0].startswith("==")]
    return rows[0], rows[1], rows[2:]


STALL = "smsp__pcsamp_warps_issue_stalled_"


def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def to_us(v, unit):
    v = float(v.replace(",", ""))
    return v * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(unit, 1)


def summarise(rep, name):
    hdr, units, data = raw(rep)
    idx = {h: i for i, h in enumerate(hdr)}
    rows = []
    for d in data:
        r = {"kernel": re.sub(r"\(.*", "", d[idx["Kernel Name"]])}
        for m, short in METRICS:
            if m not in idx:
                continue
            v, u = d[idx[m]], units[idx[m]]
            if short == "duration":
                r["duration_us"] = round(to_us(v, u), 2)
            elif short in ("dram_read", "dram_write", "l2_bytes"):
                r[short + "_MB"] = round(to_bytes(v, u) / 1e6, 3)
            else:
                r[short] = round(float(v.replace(",", "")), 3)
        dur = r["duration_us"] * 1e-6
        r["dram_GBps"] = round((r.get("dram_read_MB", 0) + r.get("dram_write_MB", 0)) * 1e6 / dur / 1e9, 1)
        if "l2_bytes_MB" in r:
            r["l2_GBps"] = round(r["l2_bytes_MB"] * 1e6 / dur / 1e9, 1)
        r["warp_execution_efficiency_pct"] = round(100 * r.get("active_threads_per_inst", 0) / 32, 1)
        # stall reasons: share of the warp-state samples (top five)
        stall = {h[len(STALL):]: float(d[idx[h]].replace(",", "")) for h in hdr if h.startswith(STALL) and "not_issued" not in h and d[idx[h]]}
        tot = sum(stall.values()) or 1.0
        for rank, (k, v) in enumerate(sorted(stall.items(), key=lambda kv: -kv[1])[:5]):
            r[f"stall{rank + 1}"] = f"{k} {100 * v / tot:.0f}%"
        rows.append(r)
    path = os.path.join(OUT, f"{tag}_{name}.csv")
    keys = list(rows[0].keys())
    with open(path, "w", newline="") as f:
        w = csv.DictWriter(f, fieldnames=keys)
        w.writeheader()
        w.writerows(rows)
    return rows


def launch_list(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    total = 0.0
    for row in csv.DictReader(lines):
        k = re.sub(r"\(.*", "", row["Kernel Name"])
        us = to_us(row["Metric Value"], row["Metric Unit"])
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += us
        total += us
    out = os.path.join(OUT, f"{tag}_launches_bench_summary.csv")
    with open(out, "w") as f:
        f.write("kernel,launches,total_us,share_pct,avg_us\n")
        for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"\"{k}\",{n},{us:.1f},{100 * us / total:.2f},{us / n:.1f}\n")
    return agg, total


os.makedirs(OUT, exist_ok=True)
g = os.path.join(ROOT, "gpurun_out")
res = {}
for name in ("extend", "shade", "shadow", "shade_mis_config4", "traversal_config5"):
    rep = os.path.join(g, f"prof_{name}_raw.csv")
    if not os.path.exists(rep):
        rep = os.path.join(g, f"prof_{name}.ncu-rep")
    if os.path.exists(rep):
        res[name] = summarise(rep, f"ncu_{name}")
        for r in res[name]:
            print(name, {k: r[k] for k in ("duration_us", "dram_GBps", "l2_GBps", "issue_active_pct", "warp_execution_efficiency_pct",
                                            "achieved_occupancy_pct", "registers", "l1_hit_pct", "l2_hit_pct") if k in r})
if os.path.exists(os.path.join(g, "launches_bench.csv")):
    import shutil
    shutil.copy(os.path.join(g, "launches_bench.csv"), os.path.join(OUT, f"{tag}_launches_bench.csv"))
    agg, total = launch_list(os.path.join(g, "launches_bench.csv"))
    for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:8]:
        print(f"{k:50s} n={n:4d} {us/1e3:9.3f} ms {100*us/total:5.1f}%")
# Per-kernel figures bench.py quotes under roofline.ncu / roofline.traffic, stamped with the hash of the kernel sources they
# were captured from (bench.py refuses them for any other sources).  The captures hold every launch of ONE bench step of
# config 2 (tools/prof_run.py: 32 spp x 1920x1080 = two waves of 33.2 M slots, 7 passes each), so "per launch" means the
# same as in bench.py's algorithmic bytes per launch.
sys.path.insert(0, ROOT)
from take_b200 import api  # noqa: E402

kernels = {}
for name in ("extend", "shade", "shadow"):
    rows = res.get(name)
    if not rows:
        continue
    dur = sum(r["duration_us"] for r in rows) * 1e-6
    insts = sum(r.get("warp_insts", 0) for r in rows)
    kernels["k_" + name] = {
        "launches_captured": len(rows),
        "duration_ms_total_under_ncu": round(dur * 1e3, 3),
        "dram_bytes_per_launch": sum(r["dram_read_MB"] + r["dram_write_MB"] for r in rows) * 1e6 / len(rows),
        "dram_gbps": round(sum(r["dram_read_MB"] + r["dram_write_MB"] for r in rows) * 1e6 / dur / 1e9, 1),
        "l2_to_sm_gbps": round(sum(r.get("l2_bytes_MB", 0) for r in rows) * 1e6 / dur / 1e9, 1),
        "issue_active_pct": round(sum(r.get("issue_active_pct", 0) * r["duration_us"] for r in rows) / (dur * 1e6), 1),
        "lanes_per_instruction": round(sum(r.get("active_threads_per_inst", 0) * r.get("warp_insts", 0) for r in rows) / max(insts, 1), 2),
        "achieved_occupancy_pct": round(sum(r.get("achieved_occupancy_pct", 0) * r["duration_us"] for r in rows) / (dur * 1e6), 1),
        "l1_hit_pct": round(sum(r.get("l1_hit_pct", 0) * r["duration_us"] for r in rows) / (dur * 1e6), 1),
        "l2_hit_pct": round(sum(r.get("l2_hit_pct", 0) * r["duration_us"] for r in rows) / (dur * 1e6), 1),
        "registers": rows[0].get("registers"),
        "rows": f"{tag}_ncu_{name}.csv",
    }
if kernels:
    head = subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
    json.dump({"tag": tag, "source_hash": api.kernel_source_hash(), "git_head_when_summarised": head,
               "workload": "tools/prof_run.py: config 2, one bench step (32 spp x 1920x1080, two waves), one_sample_mis then mis",
               "note": "ncu --set full --clock-control none; times under ncu are serialised and cold-cache: use shares, not absolutes",
               "kernels": kernels}, open(os.path.join(OUT, "ncu_metrics.json"), "w"), indent=1)
    print(json.dumps(kernels, indent=1))
