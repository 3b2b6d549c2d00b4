"""take_gpu_scene_create by phase for the BASELINE scenes: device builder (default) next to the host's SAH builder
(TAKE_DEVICE_BUILD=0).  Prints one JSON line per scene and builder.  usage: python tools/build_bench.py [c2 c3 ...]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from take_b200 import api, scenes

SCENES = {"c1": scenes.cornell_box, "c2": scenes.heightfield, "c3": scenes.ibl_scene, "c4": scenes.multi_light, "c5": scenes.instanced_spheres}
want = [a for a in sys.argv[1:] if a in SCENES] or list(SCENES)
if "--cold" in sys.argv:
    # what a process that renders one scene after another sees: every scene created ONCE, rendered (wave buffers allocated),
    # destroyed; TAKE_MEMPOOL=0 / 1 in the environment selects plain cudaMalloc or the cached pool
    for key in want:
        flat = SCENES[key]().flat()
        t0 = time.perf_counter()
        gs = api.GpuScene(flat)
        t_create = time.perf_counter() - t0
        ph = gs.create_timings()
        spp = int(os.environ.get("BB_SPP", "2"))
        _, _, st1 = gs.render_sums("mis", 5, 0, spp, seed=1, sumsq=False)
        t_first = time.perf_counter() - t0
        t2 = time.perf_counter()
        _, _, st2 = gs.render_sums("mis", 5, 0, spp, seed=1, sumsq=False)
        t_second = time.perf_counter() - t2
        t1 = time.perf_counter()
        gs.close()
        print(json.dumps({"scene": key, "mempool": os.environ.get("TAKE_MEMPOOL", "1"), "create_ms": round(1e3 * t_create, 1),
                          "create_plus_first_render_ms": round(1e3 * t_first, 1), "spp": spp, "first_render_device_ms": round(st1["ms_total"], 1),
                          "second_render_ms": round(1e3 * t_second, 1), "second_render_device_ms": round(st2["ms_total"], 1), "destroy_ms": round(1e3 * (time.perf_counter() - t1), 1),
                          "phases_ms": {k: round(v, 1) for k, v in ph.items() if k.endswith("_ms")}}), flush=True)
    sys.exit(0)
warm = api.GpuScene(scenes.cornell_box(16, 16, 1).flat()); warm.close()     # CUDA context + module load
for key in want:
    flat = SCENES[key]().flat()
    for mode in ("device", "host"):
        if mode == "host":
            os.environ["TAKE_DEVICE_BUILD"] = "0"
        else:
            os.environ.pop("TAKE_DEVICE_BUILD", None)
        best = None
        for rep in range(2):
            t0 = time.perf_counter()
            gs = api.GpuScene(flat)
            t_create = time.perf_counter() - t0
            ph = gs.create_timings()
            info = gs.info()                      # joins the background reference-order tree
            t_ready = time.perf_counter() - t0
            row = {"scene": key, "prims": flat.num_prims, "builder": mode, "create_ms": round(1e3 * t_create, 1), "ready_ms": round(1e3 * t_ready, 1),
                   "phases_ms": {k: round(v, 1) for k, v in ph.items()}, "wide_nodes": int(info["fast_nodes"]), "depth": int(info["fast_tree_depth"]),
                   "fast_tree_ms": round(info["build_ms_fast_tree"], 1), "reference_tree_ms": round(info["build_ms_reference_tree"], 1)}
            gs.close()
            if best is None or row["create_ms"] < best["create_ms"]:
                best = row
        print(json.dumps(best), flush=True)
