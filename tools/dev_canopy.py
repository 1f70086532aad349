"""Developer probe: canopy scenes, GPU 3D kernel vs CPU oracle (z-scores) + ray-caster KAT."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eradiate_b200 import kat, scenes
from eradiate_b200.kernel import mi_load_dict, render
from oracle import oracle

S = scenes.atmosphere_scene
CAN = {"lai": 2.0, "radius": 0.1, "size": (4.0, 4.0, 1.0), "padding": 1}
V3 = {"type": "mdistant", "vza": [-60.0, 0.0, 35.0], "vaa": 20.0}
cases = {
    "path_noatm": S(geometry="plane_parallel", atmosphere=None, integrator="path", canopy=CAN, sensor=V3,
                    surface={"type": "diffuse", "reflectance": 0.3}),
    "volpath_afgl": S(geometry="plane_parallel", n_layers=100, canopy=CAN, sensor=V3),
    "piecewise_afgl": S(geometry="plane_parallel", n_layers=100, integrator="piecewise_volpath", canopy=CAN, sensor=V3),
    "persp_inside": S(geometry="plane_parallel", n_layers=100, canopy=CAN,
                      sensor={"type": "perspective", "origin": [0, -7, 5], "look_at": [0, 0, 0.5], "fov": 50.0,
                              "film_resolution": (3, 2), "medium": {"type": "ref", "id": "medium_atmosphere"}}),
}
for name, kd in cases.items():
    sc = mi_load_dict(kd)
    spp = 1 << 18
    t0 = time.perf_counter()
    bmp = render(sc, sensor=0, seed=5, spp=spp)
    dt = time.perf_counter() - t0
    raw = bmp.raw
    gm = raw["sum_l"].ravel() / spp
    gv = np.maximum(raw["sum_l2"].ravel() / spp - gm**2, 0) / spp
    ospp = 1 << 14
    wl, l, l2, st = oracle.render(sc.flat.build_desc(), 0, 9, ospp)
    om = l / ospp
    ov = np.maximum(l2 / ospp - om**2, 0) / ospp
    z = (gm - om) / np.sqrt(gv + ov)
    print(f"{name:16s} gpu {np.array2string(gm, precision=5)} cpu {np.array2string(om, precision=5)} z {np.array2string(z, precision=2)}"
          f"  {bmp.stats['n_paths']/bmp.stats['device_ms']/1e3:.1f} Mpaths/s wall {dt*1e3:.1f} ms")
    print("    gpu stats", {k: round(v / bmp.stats['n_paths'], 3) for k, v in bmp.stats.items() if k.startswith(('trips', 'n_s'))},
          "cpu", {k: round(v / st['n_paths'], 3) for k, v in st.items() if k.startswith(('trips', 'n_s'))})

# ray caster
sc = mi_load_dict(cases["path_noatm"])
rng = np.random.default_rng(1)
n = 20000
o = np.stack([rng.uniform(-6, 6, n), rng.uniform(-6, 6, n), rng.uniform(0.0, 2.0, n)], axis=1)
d = rng.normal(size=(n, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
tg, ng, gg = kat.canopy_intersect(sc, o, d)
tc, nc, gc = oracle.canopy_intersect(sc.flat.build_desc(), o, d)
hit = np.isfinite(tc)
print("ray caster: hits", hit.sum(), "mismatch", (np.isfinite(tg) != hit).sum(), "max rel dt",
      np.max(np.abs(tg[hit & np.isfinite(tg)] - tc[hit & np.isfinite(tg)]) / tc[hit & np.isfinite(tg)]))
bad = np.where((np.isfinite(tg) != hit) | (hit & np.isfinite(tg) & (np.abs(tg - tc) > 1e-4 * np.maximum(tc, 1e-3))))[0]
for i in bad[:12]:
    zc = o[i, 2] + tc[i] * d[i, 2] if np.isfinite(tc[i]) else np.nan
    zg = o[i, 2] + tg[i] * d[i, 2] if np.isfinite(tg[i]) else np.nan
    print("  o", np.round(o[i], 3), "d", np.round(d[i], 3), "t_cpu", tc[i], "t_gpu", tg[i], "z_cpu", zc, "z_gpu", zg)
sc = mi_load_dict(cases["path_noatm"])
spp = 1 << 22
bmp = render(sc, sensor=0, seed=15, spp=spp)
gm = bmp.raw["sum_l"].ravel() / spp
ospp = 1 << 17
wl, l, l2, st = oracle.render(sc.flat.build_desc(), 0, 19, ospp)
om = l / ospp; ov = np.maximum(l2 / ospp - om**2, 0) / ospp
print("hi-spp noatm: gpu", gm, "cpu", om, "z", (gm - om) / np.sqrt(ov), "Mpaths/s", bmp.stats['n_paths']/bmp.stats['device_ms']/1e3)
