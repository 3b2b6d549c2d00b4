"""Generate tests/golden/*.npz + *.takescene from the UNMODIFIED reference (oracle/_ref/libtake_ref.so).

Run in the build container (needs /root/reference to have been compiled by `make -C oracle ref`):
    python tests/golden/make_golden.py
The reference ships no tests, golden vectors or scenes of its own, so these vectors -- the
reference's own outputs on our generated scenes -- are what pins the oracle on machines without the reference.
Each .npz holds, for one small scene: the reference's BVH link array, (primitive id, t, hit record) of 1024 jittered
primary rays and of one random secondary ray per hit, occlusion flags of finite segments, and the per-pixel radiance
sums / sums of squares of 2 samples per pixel for each of the three live integrators (seed 2024, max_depth 5).
"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import bindings as ob  # noqa: E402
from take_b200 import scenes  # noqa: E402

SCENES = {
    "cornell": lambda: scenes.cornell_box(32, 32, 2),
    "cornell_mixed": lambda: scenes.cornell_box(32, 32, 2, materials="mixed"),
    "multi_light": lambda: scenes.multi_light(40, 24, 2, n_side=4),
    "heightfield": lambda: scenes.heightfield(24, 40, 24, 2),
    "textured": lambda: scenes.textured_room(32, 32, 2),
    "spheres": lambda: scenes.sphere_room(32, 32, 2),
    "cornell_stubs": lambda: scenes.cornell_stubs(32, 32, 2),   # Disney stub BSDFs 7, 8, 10, 11 + a point emitter
}
SEED = 2024


def main():
    R = ob.RefLib()
    O = ob.OracleLib()
    assert R.rng_selfcheck(SEED, 5, 9) == 0
    for name, make in SCENES.items():
        b = make()
        flat = b.flat()
        with tempfile.TemporaryDirectory() as d:
            rs = R.load(b.write(d))
            rs.dump(os.path.join(d, "ref.takescene"))
            from take_b200.sceneio import FlatScene
            assert flat.same_as(FlatScene.load(os.path.join(d, "ref.takescene"))) == [], name
            flat.save(os.path.join(HERE, f"{name}.takescene"))
            os_ = O.load(flat)  # only used to generate camera rays (checked against the reference below)
            H, W = flat.height, flat.width
            rng = np.random.default_rng(1)
            pix = rng.choice(H * W, min(1024, H * W), replace=False)
            rays = os_.primary_rays(pix % W, pix // W, seed=SEED)
            prim, t, rec = rs.intersect(rays, records=True)
            sec = ob.secondary_rays(rays, t, prim, seed=3)
            prim2, t2, rec2 = rs.intersect(sec, records=True)
            seg = sec.copy()
            seg[:, 7] = rng.uniform(0.05, 2.0, len(seg)) * np.abs(flat.positions).max() * 0.3
            occ = rs.occluded(seg)
            _, links = rs.bvh()
            out = dict(rays=rays, prim=prim, t=t, rec=rec, sec=sec, prim2=prim2, t2=t2, rec2=rec2, seg=seg, occ=occ,
                       bvh_links=links, bvh_root=np.int32(rs.root), seed=np.int64(SEED))
            for integ in ob.INTEGRATORS:
                s, s2 = rs.render(integ, 5, 0, 2, seed=SEED)
                out[f"sum_{integ}"], out[f"sumsq_{integ}"] = s, s2
            np.savez_compressed(os.path.join(HERE, f"{name}.npz"), **out)
            print(name, "prims", flat.num_prims, "hits", int((prim >= 0).sum()), "/", len(prim))


if __name__ == "__main__":
    main()
