"""Per-source-line view of an `ncu --set full --import-source on` capture (development aid).

ncu's `--page source --csv` lists SASS instructions with stall samples and executed-instruction counts; nvdisasm gives
the source line of every SASS offset (the library is compiled with -lineinfo).  This joins the two and prints the
hottest source lines of one kernel launch.

usage: python tools/ncu_lines.py <report.ncu-rep> <mangled-kernel-substring> [launch-index] [top-n]
"""
import collections, csv, io, os, re, subprocess, sys, tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, pat = sys.argv[1], sys.argv[2]
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
topn = int(sys.argv[4]) if len(sys.argv) > 4 else 40
lib = os.environ.get("TAKE_GPU_LIB", os.path.join(ROOT, "take_b200", "libtake_gpu.so"))

tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.startswith("take_gpu.") and f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
line_of = {}
on, cur = False, ("?", 0)
for l in dis.splitlines():
    if l.startswith("//---") and ".text." in l:
        on = pat in l
        continue
    if not on:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*)", l)
    if m:
        line_of[int(m.group(1), 16)] = (cur, m.group(2).strip())

out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
sections, cur_rows, hdr = [], None, None
for row in csv.reader(io.StringIO(out)):
    if not row:
        continue
    if row[0] == "Kernel Name":
        cur_rows = []
        sections.append((row[1], cur_rows))
    elif row[0] == "Address":
        hdr = row
    elif cur_rows is not None and row[0].startswith("0x"):
        cur_rows.append(row)
name, rows = sections[which]
ix = {h: i for i, h in enumerate(hdr)}
base = int(rows[0][0], 16)
agg = collections.defaultdict(lambda: [0, 0, 0, collections.Counter()])
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot_s = tot_i = tot_t = 0
for r in rows:
    off = int(r[0], 16) - base
    key, _ = line_of.get(off, (("?", 0), ""))
    s, i, t = int(r[ix["# Samples"]]), int(r[ix["Instructions Executed"]]), int(r[ix["Thread Instructions Executed"]])
    a = agg[key]
    a[0] += s; a[1] += i; a[2] += t
    for c in stall_cols:
        v = int(r[ix[c]] or 0)
        if v:
            a[3][c] += v
    tot_s += s; tot_i += i; tot_t += t
print(f"{name}\n  samples {tot_s}  warp-insts {tot_i}  thread-insts {tot_t}  avg active lanes {tot_t / max(tot_i, 1):.1f}")
src_cache = {}
def src(key):
    f, n = key
    for d in ("take_b200/csrc", "include"):
        p = os.path.join(ROOT, d, f)
        if os.path.exists(p):
            if p not in src_cache:
                src_cache[p] = open(p).read().splitlines()
            return src_cache[p][n - 1].strip()[:90] if 0 < n <= len(src_cache[p]) else ""
    return ""
print(f"{'file:line':28s} {'samp%':>6s} {'inst%':>6s} {'lanes':>5s}  top stalls | source")
for key, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:topn]:
    st = ",".join(f"{k[6:]}:{v}" for k, v in a[3].most_common(3))
    print(f"{key[0] + ':' + str(key[1]):28s} {100 * a[0] / tot_s:6.2f} {100 * a[1] / tot_i:6.2f} {a[2] / max(a[1], 1):5.1f}  {st} | {src(key)}")
