"""Write profiles/README.md from the committed summaries (run after tools/ncu_summary.py).  usage: python tools/profiles_readme.py [tag=r02]"""
import csv, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TAG = sys.argv[1] if len(sys.argv) > 1 else "r02"
P = lambda n: os.path.join(ROOT, "profiles", n)
rows = lambda n: list(csv.DictReader(open(P(n)))) if os.path.exists(P(n)) else []
ext, sh, sw = rows(f"{TAG}_ncu_extend.csv"), rows(f"{TAG}_ncu_shade.csv"), rows(f"{TAG}_ncu_shadow.csv")
c4, c5 = rows(f"{TAG}_ncu_shade_mis_config4.csv"), rows(f"{TAG}_ncu_traversal_config5.csv")
b = json.load(open(P(f"{TAG}_bench_n1.json")))
m = json.load(open(P("ncu_metrics.json")))


def short(k):
    return k.replace("void ", "").replace("take::", "").strip()


def row(r, label):
    return (f"| {label} | `{short(r['kernel'])}` | {float(r['duration_us']):.0f} | {r['dram_GBps']} | {r.get('l2_GBps', '-')} | {float(r.get('l1_pct_of_peak', 0)):.0f} | "
            f"{float(r['issue_active_pct']):.0f} | {float(r['active_threads_per_inst']):.1f} | {float(r['achieved_occupancy_pct']):.0f} | {float(r['registers']):.0f} | "
            f"{float(r['l1_hit_pct']):.0f} / {float(r['l2_hit_pct']):.0f} | {r.get('stall1', '')}, {r.get('stall2', '')}, {r.get('stall3', '')} |")


HDR = ("| launch | kernel | time (us) | DRAM GB/s | L2->SM GB/s | L1TEX % of peak | issue slots used % | lanes per instruction | achieved occupancy % | regs | "
       "L1 / L2 hit % | top stalls (share of samples) |\n|---|---|---|---|---|---|---|---|---|---|---|---|")
L = [f"# profiles/ — round 2 (`{TAG}_*`; `r01_*` are round 1's files, kept for comparison)\n",
     "All captures: B200, `ncu --set full --clock-control none --import-source on`, taken by ONE command (`tools/gpu_round.sh`, run under\n"
     "`gpurun`) from the kernels of the commit named in `ncu_metrics.json`; the raw metric pages are exported on the box and the\n"
     "`.ncu-rep` files deleted there (gpurun carries 64 MiB back).  `tools/ncu_summary.py r02` produced the CSVs and\n"
     "`ncu_metrics.json`, `tools/profiles_readme.py` this file.  Times under ncu are serialised and cold-cache: use the shares.\n",
     "| file | what |\n|---|---|",
     f"| `{TAG}_bench_n1.json`, `{TAG}_bench_reference_n1.json` | the `bench.py` line (ours) and the reference arm, same box |",
     f"| `{TAG}_scale_n1/2/4/8.json` | `bench.py --gpus N` under torchrun on 1, 2, 4, 8 B200 of one box: weak scaling of the headline, strong scaling of every config (`scenes`), in-run `checks` |",
     f"| `{TAG}_launches_bench.csv`, `{TAG}_launches_bench_summary.csv` | `ncu --metrics gpu__time_duration.sum` launch list of `bench.py --steps 2 --warmup 3 --no-cpu-baseline --scenes ''` (first 300 launches) and its per-kernel totals |",
     f"| `{TAG}_ncu_extend.csv`, `{TAG}_ncu_shade.csv`, `{TAG}_ncu_shadow.csv` | key metrics + stall shares of EVERY launch of one bench step of config 2 (`tools/prof_run.py`: 32 spp x 1920x1080 = two waves of 33.2 M slots, 7 passes each -> 14 extend and 14 shade launches; the shadow kernel from a 16-spp multi-sample wave) |",
     "| `ncu_metrics.json` | per-kernel aggregates of those captures (DRAM bytes per launch, L2->SM GB/s, issue-active %, lanes per instruction), stamped with a hash of the kernel sources; `bench.py` quotes them under `roofline.ncu` / `roofline.traffic` only when the hash matches the sources it runs |",
     f"| `{TAG}_ncu_shade_mis_config4.csv`, `{TAG}_ncu_traversal_config5.csv` | the same for the multi-sample shade kernel on the config-4 scene (`--scene=multi_light --integrator=mis --spp=16`, passes 0-6 of a wave) and the traversal kernels on the 10 M-triangle scene (`--scene=instanced --spp=4`: `k_extend_primary`, then `k_extend_refill` / `k_shadow_refill` alternating) |",
     f"| `{TAG}_lines_extend_pass2_one_ray_per_thread.txt`, `{TAG}_lines_extend_pass1_lane_refill.txt`, `{TAG}_lines_shade_mis_config4_pass1.txt`, `{TAG}_lines_extend_primary.txt` | `tools/lines_round.sh` / `tools/ncu_lines.py`: stall samples, instruction share and active lanes PER SOURCE LINE of a bounce pass of config 2 traced one ray per thread (`TAKE_REFILL=0`: 9 of 32 lanes) and with lane refill (17.5 lanes), of the multi-sample shade kernel on the config-4 scene, and of the camera-ray packet kernel (the four box tests per lane and node are 36 % of its instructions) |",
     f"| `{TAG}_prof_run_counts.txt` | box / leaf test counts of the profiled step (instrumented kernels) |",
     f"| `{TAG}_build_bench.jsonl` | `tools/build_bench.py`: `take_gpu_scene_create` per config with the device builder and with the host builders, phase by phase |",
     f"| `{TAG}_device_vs_host_tree.jsonl` | `tools/tune.py`: a 32-spp step on the device-built and on the host-built tree (three scenes) |",
     f"| `{TAG}_mesh_load_1M.json`, `{TAG}_mesh_load_10M.json` | `tools/mesh_load_bench.py --ref`: PLY -> flat scene arrays through `take_gpu_builder_add_ply` and through the unmodified reference parser, arrays compared |\n",
     "## Share of a step (launch list vs. bench.py's own CUDA-event stage times)\n",
     "| kernel | launches | share under ncu | stage in `roofline.stage_share` (bench.py) |\n|---|---|---|---|"]
ss = b["roofline"]["stage_share"]
stage = {"k_extend": "extend", "k_shade": "shade", "k_accumulate": "other", "k_scatter": "sort"}
for r in rows(f"{TAG}_launches_bench_summary.csv")[:8]:
    key = [v for k, v in stage.items() if k in r["kernel"]]
    L.append(f"| `{short(r['kernel'])}` | {r['launches']} | {float(r['share_pct']):.1f} % | {key[0] + ' ' + format(100 * ss[key[0]], '.1f') + ' % (all kernels of the stage)' if key else 'scene creation'} |")
L += ["\nThe dominant kernels are the extend kernels (4-wide BVH traversal + FP64 leaf tests) in both views.\n",
      "## Extend kernels — the 14 launches of one bench step (wave 0: the sky half of the frame, wave 1: the heavy half)\n", HDR]
L += [row(r, str(i)) for i, r in enumerate(ext)]
L += ["\n## `k_shade<one_sample_mis>` — the 14 launches of the same step\n", HDR] + [row(r, str(i)) for i, r in enumerate(sh)]
L += ["\n## `k_shadow_refill` (multi-sample MIS wave on config 2)\n", HDR] + [row(r, str(i)) for i, r in enumerate(sw)]
if c4:
    L += ["\n## `k_shade<mis>` on the config-4 scene (400 emitters)\n", HDR] + [row(r, str(i)) for i, r in enumerate(c4)]
if c5:
    L += ["\n## Traversal kernels on the 10 M-triangle scene (config 5)\n", HDR] + [row(r, str(i)) for i, r in enumerate(c5)]
k = m["kernels"]
r = b["roofline"]
prim = [x for x in ext if "primary" in x["kernel"]]
ref1 = [x for x in ext if "refill" in x["kernel"]]
L.append(f"""
## Reading

* The extend kernels are **not HBM-bound**: {k['k_extend']['dram_gbps']} GB/s of DRAM traffic over the step ({100 * k['k_extend']['dram_gbps'] / r['peak']:.0f} % of the measured
  {r['peak']:.0f} GB/s copy bandwidth), {k['k_extend']['l2_to_sm_gbps']} GB/s from L2; measured DRAM traffic per launch ({k['k_extend']['dram_bytes_per_launch'] / 1e6:.0f} MB) is
  {r['bytes_per_launch'] / k['k_extend']['dram_bytes_per_launch']:.0f}x *below* the algorithmic bytes ({r['bytes_per_launch'] / 1e6:.0f} MB per launch by the SURVEY.md 8(d) counting
  rule) because the 130 MB of tree and leaf records live in L1/L2.  `roofline.achieved` = algorithmic bytes / time =
  {r['achieved']:.0f} GB/s = {r['frac']:.2f} of the HBM peak -- a statement about the counting rule, not about DRAM.
* Camera rays (`k_extend_primary`, packets): {float(prim[0]['active_threads_per_inst']):.1f} of 32 lanes, {float(prim[0]['issue_active_pct']):.0f} % of the issue slots -- issue-bound.
* Bounce rays (`k_extend_refill`, lanes refilled): {float(ref1[0]['active_threads_per_inst']):.1f} lanes per instruction in the first bounce pass
  (8.8 with one ray per thread, `r01_ncu_extend.csv`), L1TEX at {float(ref1[0].get('l1_pct_of_peak', 0)):.0f} % of its peak, {float(ref1[0]['issue_active_pct']):.0f} % of the issue slots, top stall
  {ref1[0].get('stall1')}: latency-bound with the L1 close behind.  Whole step: {k['k_extend']['lanes_per_instruction']} lanes per instruction (15.3 before).
* Ray-box / ray-triangle test rate against FP32 peak: {r['box_tests_per_ray']:.1f} box + {r['tri_tests_per_ray']:.2f} leaf tests per ray at
  {r['extend_grays_per_s']:.2f} Grays/s = {r['test_rate_tflops']:.2f} TFLOP/s by the 27 / 60 flop counting rule = {100 * r['test_rate_frac_fp32']:.1f} % of the
  {r['fp32_peak_tflops']:.1f} TFLOP/s FP32 peak (the leaf tests actually run in FP64).
* `k_shade` moves {k['k_shade']['dram_gbps'] / 1e3:.1f} TB/s through DRAM over the step ({100 * k['k_shade']['dram_gbps'] / r['peak']:.0f} % of peak) at {k['k_shade']['achieved_occupancy_pct']:.0f} % occupancy (128 registers of FP64
  state), {k['k_shade']['issue_active_pct']} % of the issue slots; top stalls `long_scoreboard` (gathered records) then `wait` -- latency-bound.
""")
open(P("README.md"), "w").write("\n".join(L))
print("profiles/README.md written")
