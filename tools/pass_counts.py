"""Per-pass box / leaf test counts of k_extend on config 2 (development aid): differences of instrumented renders with
max_depth = -1, 0, 1, 2."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from take_b200 import api, scenes
gs = api.GpuScene(scenes.heightfield().flat())
prev = None
for md in (-1, 0, 1, 2, 5):
    _, _, st = gs.render_sums("one_sample_mis", md, 0, 2, seed=1, sumsq=False, flags=api.RENDER_COUNT_TESTS)
    cur = (st["extend_rays"], st["box_tests"], st["tri_tests"])
    if prev:
        dr, db, dt = (c - p for c, p in zip(cur, prev))
        print(f"passes up to max_depth {md}: +{dr} rays, {db/max(dr,1):.1f} box tests/ray, {dt/max(dr,1):.2f} leaf tests/ray")
    else:
        print(f"pass 0: {cur[0]} rays, {cur[1]/cur[0]:.1f} box tests/ray, {cur[2]/cur[0]:.2f} leaf tests/ray")
    prev = cur
