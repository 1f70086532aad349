import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from eradiate_b200 import scenes
from eradiate_b200.kernel import mi_load_dict, render
from oracle import oracle
from tests.util import stats_from_sums

for geom in ("plane_parallel", "spherical_shell"):
    d = scenes.atmosphere_scene(geometry=geom, n_layers=100, sensor={"type": "mdistant", "vza": [60., 80., 85., 88., 89.5], "vaa": 0.0})
    sc = mi_load_dict(d)
    spp_o = 1 << 16
    wl, l, l2, st = oracle.render(sc.flat.build_desc(), 0, 3, spp_o)
    mo, vo = stats_from_sums(l, l2, spp_o)
    spp = 1 << 20
    bmp = render(sc, 0, 5, spp)
    mg, vg = stats_from_sums(bmp.raw["sum_l"].ravel(), bmp.raw["sum_l2"].ravel(), spp)
    print(geom, "cpu", mo, "\n   gpu", mg, "\n   z", (mg - mo) / np.sqrt(vo + vg), "\n  K cpu", (st["trips_main"]+st["trips_nee"])/st["n_paths"], "K gpu", (bmp.stats["trips_main"]+bmp.stats["trips_nee"])/bmp.stats["n_paths"])
