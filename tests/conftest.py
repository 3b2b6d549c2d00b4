import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import bindings as ob  # noqa: E402  (tests are allowed to use the checkers)
from take_b200 import api, scenes  # noqa: E402


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _ensure_built():
    if not os.path.exists(ob.ORACLE_SO):
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "port"], check=True)
    if not os.path.exists(api.LIB_PATH):
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "take_b200", "csrc")], check=True)


@pytest.fixture(scope="session")
def oracle_lib():
    _ensure_built()
    return ob.OracleLib()


@pytest.fixture(scope="session")
def ref_lib():
    if not ob.have_ref():
        pytest.skip("oracle/_ref/libtake_ref.so not built (needs /root/reference)")
    return ob.RefLib()


@pytest.fixture(scope="session")
def gpu_lib():
    _ensure_built()
    return api.load_library()


SMALL_SCENES = {
    "cornell": lambda: scenes.cornell_box(48, 48, 4),
    "cornell_mixed": lambda: scenes.cornell_box(48, 48, 4, materials="mixed"),
    "multi_light": lambda: scenes.multi_light(64, 40, 4, n_side=5),
    "heightfield": lambda: scenes.heightfield(48, 64, 36, 4),
    "textured": lambda: scenes.textured_room(48, 48, 4),
    "spheres": lambda: scenes.sphere_room(48, 48, 4),
    "cornell_stubs": lambda: scenes.cornell_stubs(48, 48, 4),   # Disney stub BSDFs 7, 8, 10, 11 + a point emitter
}


@pytest.fixture(scope="session", params=sorted(SMALL_SCENES))
def small_scene(request):
    """(name, SceneBuilder-or-None, FlatScene) for each small test scene."""
    b = SMALL_SCENES[request.param]()
    flat = b if not hasattr(b, "flat") else b.flat()
    return request.param, (b if hasattr(b, "write") else None), flat


def all_pixel_rays(oscene, seed=7, jitter=True):
    H, W = oscene.height, oscene.width
    py, px = np.mgrid[0:H, 0:W]
    return oscene.primary_rays(px.ravel(), py.ravel(), seed=seed, jitter=jitter)
