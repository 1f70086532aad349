"""Developer probe: one C4-like canopy render (for ncu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eradiate_b200 import scenes
from eradiate_b200.kernel import mi_load_dict, render
kd = scenes.config_c4(spp=1 << 12, lai=3.0, radius=0.1, size=(10.0, 10.0, 2.0), padding=2, n_vza=32, film=(8, 8))
sc = mi_load_dict(kd)
bmp = render(sc, sensor=0, seed=3, spp=1 << 15)
print(bmp.stats)
