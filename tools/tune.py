"""A/B timing of kernel variants on the config-2 scene (development aid).  One process per variant (env selects it)."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

def child():
    from take_b200 import api, scenes
    integ = os.environ.get("TUNE_INTEGRATOR", "one_sample_mis")
    spp = int(os.environ.get("TUNE_SPP", "16"))
    flat = scenes.build(os.environ.get("TUNE_SCENE", "heightfield")).flat()
    gs = api.GpuScene(flat)
    best = None
    for rep in range(4):
        s, s2, st = gs.render_sums(integ, 5, 0, spp, seed=1, flags=api.RENDER_STAGE_TIMES if rep == 3 else 0)
        if rep < 3 and (best is None or st["ms_total"] < best): best = st["ms_total"]
    rays = st["extend_rays"] + st["shadow_rays"]
    print(json.dumps(dict(tag=os.environ.get("TUNE_TAG"), ms_best=round(best, 3), mrays=round(rays / best / 1e3, 1),
                          ext_ms=round(st["ms_extend"], 3), ext_grays=round(st["extend_rays"] / st["ms_extend"] / 1e6, 3),
                          shade_ms=round(st["ms_shade"], 3), shadow_ms=round(st["ms_shadow"], 3), gen_ms=round(st["ms_generate"], 3),
                          sort_ms=round(st["ms_sort"], 3), other_ms=round(st["ms_other"], 3), sum=float(s.sum()))))
    gs.close()

if os.environ.get("TUNE_CHILD"):
    child()
    sys.exit(0)

variants = []
for a in sys.argv[1:]:
    # tag:ENV=VAL,ENV=VAL
    tag, _, envs = a.partition(":")
    env = dict(e.split("=", 1) for e in envs.split(",") if e)
    variants.append((tag, env))
for tag, env in variants:
    e = dict(os.environ, TUNE_CHILD="1", TUNE_TAG=tag, **env)
    if "LIB" in env:
        e["TAKE_GPU_LIB"] = os.path.join(ROOT, "take_b200", env["LIB"])
    r = subprocess.run([sys.executable, __file__], env=e, capture_output=True, text=True)
    print(r.stdout.strip() or ("FAILED " + tag + " " + r.stderr[-400:]))
    if "[take_gpu]" in r.stderr:
        print("\n".join(l for l in r.stderr.splitlines() if l.startswith("[take_gpu]")))
