import os, sys, json, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eradiate_b200 import scenes
from eradiate_b200.kernel import mi_load_dict, render
from oracle import oracle
S = scenes.atmosphere_scene
OC = {"type": "ocean_legacy", "wavelength": 865.0, "wind_speed": 5.0, "wind_direction": 30.0, "chlorinity": 19.0, "pigmentation": 0.3, "shadowing": True}
SEN = {"type": "mdistant", "vza": np.linspace(-60.0, 60.0, 4), "vaa": 40.0}
def c5(geometry):
    return scenes.config_c5(spp=16, n_vza=4, w_nm=865.0, n_layers=120, geometry=geometry)
cases = {
    "A c5 plane-parallel": c5("plane_parallel"),
    "B sph ocean rayleigh_pol only": S(geometry="spherical_shell", n_layers=120, w_nm=865.0, phase={"type": "rayleigh_polarized", "depolarization": 0.0279}, stokes=True, sza=35.0, surface=dict(OC), sensor=dict(SEN)),
    "C sph ocean wind_dir 0": S(geometry="spherical_shell", n_layers=120, w_nm=550.0, phase={"type": "rayleigh_polarized"}, stokes=True, sza=35.0, surface=dict(OC, wavelength=550.0, wind_direction=0.0), sensor=dict(SEN)),
    "D sph ocean no atmosphere pol": S(geometry="spherical_shell", atmosphere=None, stokes=True, sza=35.0, surface=dict(OC), sensor=dict(SEN)),
    "E pp ocean rayleigh_pol 550 thick": S(geometry="plane_parallel", n_layers=120, w_nm=400.0, phase={"type": "rayleigh_polarized"}, stokes=True, sza=35.0, surface=dict(OC, wavelength=400.0), sensor=dict(SEN)),
    "F sph ocean rayleigh_pol 400 thick": S(geometry="spherical_shell", n_layers=120, w_nm=400.0, phase={"type": "rayleigh_polarized"}, stokes=True, sza=35.0, surface=dict(OC, wavelength=400.0), sensor=dict(SEN)),
}
for name, kd in cases.items():
    sc = mi_load_dict(kd)
    if "no atmosphere" in name: sc._force_polarized = True
    spp = 1 << 21
    bmp = render(sc, seed=3, spp=spp)
    m = bmp.raw["sum_l"].ravel() / spp
    v = np.maximum(bmp.raw["sum_l2"].ravel() / spp - m * m, 0) / spp
    ospp = 1 << 17
    wl, l, l2, st4, st = oracle.render_stokes(sc.flat.build_desc(), 0, 5, ospp)
    om = l / ospp; ov = np.maximum(l2 / ospp - om * om, 0) / ospp
    z = (m - om) / np.sqrt(v + ov)
    sg = bmp.raw["sum_stokes"].reshape(4, -1) / spp
    print(name, "\n   gpu I", np.round(m, 5), "cpu I", np.round(om, 5), "z", np.round(z, 2),
          "\n   gpu Q", np.round(sg[1], 5), "cpu Q", np.round(st4[1] / ospp, 5), "gpu U", np.round(sg[2], 5), "cpu U", np.round(st4[2] / ospp, 5))
