import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eradiate_b200 import scenes
from eradiate_b200.kernel import mi_load_dict, render
sc = mi_load_dict(scenes.config_c5())
render(sc, seed=1, spp=1 << 18)
for r in range(2):
    bmp = render(sc, seed=2 + r, spp=1 << 25)
    st = bmp.stats
    print(f"C5 band: {st['n_paths']/st['device_ms']/1e3:.1f} Mpaths/s, K={(st['trips_main']+st['trips_nee'])/st['n_paths']:.1f}")
