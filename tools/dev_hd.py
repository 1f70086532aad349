import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from eradiate_b200 import scenes
from eradiate_b200.kernel import mi_load_dict, render
from oracle import oracle
from tests.util import stats_from_sums

def cmp(name, d, spp_o=1<<16, spp=1<<20):
    sc = mi_load_dict(d)
    wl, l, l2, st = oracle.render(sc.flat.build_desc(), 0, 3, spp_o)
    mo, vo = stats_from_sums(l, l2, spp_o)
    bmp = render(sc, 0, 5, spp)
    mg, vg = stats_from_sums(bmp.raw["sum_l"].ravel(), bmp.raw["sum_l2"].ravel(), spp)
    print(f"{name}\n   cpu {mo}\n   gpu {mg}\n   z {(mg - mo) / np.sqrt(vo + vg + 1e-30)}\n   ratio {mg/mo}")
    print("   K cpu", st["trips_main"]/st["n_paths"], st["trips_nee"]/st["n_paths"], st["n_scatter"]/st["n_paths"], st["n_surface"]/st["n_paths"],
          "K gpu", bmp.stats["trips_main"]/bmp.stats["n_paths"], bmp.stats["trips_nee"]/bmp.stats["n_paths"], bmp.stats["n_scatter"]/bmp.stats["n_paths"], bmp.stats["n_surface"]/bmp.stats["n_paths"])

S = scenes.atmosphere_scene
cmp("pp mdistant vaa=90", S(geometry="plane_parallel", n_layers=100, sensor={"type": "mdistant", "vza": [20., 50., 70.], "vaa": 90.0}))
cmp("pp hdistant 1x1 rpv", S(geometry="plane_parallel", n_layers=100, sensor={"type": "hdistant", "film_resolution": (1, 1)}))
cmp("pp hdistant 1x1 lambert", S(geometry="plane_parallel", n_layers=100, surface={"type": "diffuse", "reflectance": 0.3}, sensor={"type": "hdistant", "film_resolution": (1, 1)}))
cmp("pp hdistant 2x2 noatm rpv", S(geometry="plane_parallel", atmosphere=None, sensor={"type": "hdistant", "film_resolution": (2, 2)}))
cmp("pp hdistant 1x1 black surface", S(geometry="plane_parallel", n_layers=100, surface={"type": "diffuse", "reflectance": 0.0}, sensor={"type": "hdistant", "film_resolution": (1, 1)}))
cmp("sph hdistant 1x1", S(geometry="spherical_shell", n_layers=100, sensor={"type": "hdistant", "film_resolution": (1, 1)}))
cmp("sph mdistant notarget", S(geometry="spherical_shell", n_layers=100, sensor={"type": "mdistant", "vza": [0., 40.], "vaa": 0.0, "target": None}))
