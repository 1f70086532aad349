#!/usr/bin/env python
"""
Small renders through every kernel family, for `compute-sanitizer` (racecheck / memcheck / synccheck):

    compute-sanitizer --tool racecheck python tools/sanitize_probe.py
    compute-sanitizer --tool memcheck  python tools/sanitize_probe.py

The warp-private shared-memory pools of the pool kernel rely on __syncwarp ordering between the compaction
list, the record loads and the record stores: racecheck is the tool that would see a missing one.  Sample counts
are tiny (the tools slow kernels down 10-100x); every instance family is covered: scalar pool (spherical,
plane-parallel), BANDS, POL, POL+BANDS, PW, GEN (mpdistant, astroobject), COLL flush, the register kernel and the
3D canopy kernel.  Prints one line per scene with the film mean so that a run is visibly doing work.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

from eradiate_b200.kernel import mi_load_dict, render  # noqa: E402
from tests.scene_battery import battery  # noqa: E402

SCENES = [
    ("c2_afgl_rpv_spherical", 256), ("afgl_rpv_pp", 256), ("c3_afgl_aerosol_tab_hdistant", 32),
    ("polarized_rayleigh_pp", 128), ("c5_polarized_ocean_aerosol_reduced", 64), ("piecewise_afgl_rpv_pp", 256),
    ("polarized_piecewise_rayleigh_pp", 128), ("mpdistant_spherical", 128), ("astro_sun_aerosol_tab_spherical", 32),
    ("polarized_astro_mishchenko_pp", 64), ("ocean_grasp_spherical", 128), ("canopy_volpath_afgl_rpv_pp", 64),
    ("canopy_perspective_inside_pp", 64), ("canopy_abstract_trees_pp", 64),
]


def main():
    bat = battery()
    for name, spp in SCENES:
        img = np.array(render(mi_load_dict(bat[name]), sensor=0, seed=3, spp=spp))[..., 0]
        print(f"{name:44s} spp {spp:4d}  mean {img.mean():.6g}  finite {bool(np.isfinite(img).all())}", flush=True)
    os.environ["ERTB_KERNEL"] = "legacy"
    for name, spp in SCENES[:3]:
        img = np.array(render(mi_load_dict(bat[name]), sensor=0, seed=3, spp=spp))[..., 0]
        print(f"{name:44s} spp {spp:4d}  mean {img.mean():.6g}  (register kernel)", flush=True)


if __name__ == "__main__":
    main()
