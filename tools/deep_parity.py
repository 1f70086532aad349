#!/usr/bin/env python
"""
Deep parity probe: the newer kernels (3D canopy kernel, GEN / BANDS pool instances) against the CPU oracle at
sample counts 8-16x those of the committed fixtures, to expose biases the regular battery is too noisy to see.
Prints per-scene z-scores; exits 1 if any |z| > 4. Needs a GPU; takes a few minutes of host time.
"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eradiate_b200.kernel import mi_load_dict, render
from oracle import oracle
from tests.scene_battery import battery

NAMES = ["c2_afgl_rpv_spherical", "afgl_rpv_pp", "thick_isotropic_pp", "ocean_pp", "polarized_rayleigh_pp", "piecewise_afgl_rpv_pp",
         "canopy_abstract_trees_pp", "canopy_volpath_afgl_rpv_pp", "canopy_piecewise_aerosol_pp", "canopy_perspective_inside_pp",
         "canopy_path_no_atmosphere", "c4_canopy_afgl_rpv_reduced", "central_patch_canopy_mpdistant_pp",
         "mradiancemeter_sky_and_nadir_spherical", "mradiancemeter_piecewise_aerosol_pp", "mpdistant_spherical",
         "c3_afgl_aerosol_tab_hdistant", "aerosol_tab_irregular_spherical"]
# SURVEY 8f-4 plugins added last (glint family, mqdiffuse, multiphase); DEEP_NAMES=new selects only these
NEW = ["ocean_mishchenko_pp", "ocean_grasp_spherical", "maignan_pp", "polarized_mishchenko_pp", "polarized_grasp_spherical",
       "polarized_maignan_pp", "mqdiffuse_pp", "mqdiffuse_spherical_thick", "polarized_mqdiffuse_pp",
       "multiphase_three_components_pp", "astro_wide_disc_afgl_rpv_pp", "astro_sun_aerosol_tab_spherical",
       "astro_piecewise_ocean_grasp_pp", "astro_direct_beam_from_ground_spherical", "polarized_astro_mishchenko_pp"]
sel = os.environ.get("DEEP_NAMES", "")
NAMES = NEW if sel == "new" else (sel.split(",") if sel else NAMES + NEW)
bad = 0
B = battery()
for name in NAMES:
    sc = mi_load_dict(B[name])
    d = sc.flat.build_desc()
    heavy = name.startswith(("c3_", "aerosol", "astro_sun_aerosol"))
    ospp = 1 << ((17 if heavy else 20) + int(os.environ.get("DEEP", "0")))
    t0 = time.perf_counter()
    ostokes = None
    if d.polarized:
        wl, l, l2, ostokes, st = oracle.render_stokes(d, 0, 77, ospp)
    else:
        wl, l, l2, st = oracle.render(d, 0, 77, ospp)
    om = l / ospp
    ov = np.maximum(l2 / ospp - om * om, 0) / ospp
    t1 = time.perf_counter()
    gspp = 1 << (23 + int(os.environ.get("DEEP", "0")))
    bmp = render(sc, seed=91, spp=gspp)
    gm = bmp.raw["sum_l"].ravel() / gspp
    gv = np.maximum(bmp.raw["sum_l2"].ravel() / gspp - gm * gm, 0) / gspp
    z = (gm - om) / np.sqrt(gv + ov + 1e-300)
    rel = np.max(np.abs(gm - om) / np.maximum(om, 1e-12))
    if ostokes is not None:  # Q, U, V: per-sample |Q| <= I, so the second moment of I bounds their variance
        gs = np.asarray(bmp.raw["sum_stokes"]).reshape(4, -1) / gspp
        os_ = np.asarray(ostokes).reshape(4, -1) / ospp
        bound = np.sqrt(bmp.raw["sum_l2"].ravel() / gspp / gspp + l2 / ospp / ospp)
        zq = np.abs(gs[1:] - os_[1:]) / bound
        z = np.concatenate([z, zq.ravel()])
    flag = "" if np.all(np.abs(z) <= 4.0) else "   <-- CHECK"
    bad += flag != ""
    print(f"{name:44s} |z|max {np.abs(z).max():4.2f}  max rel diff {rel:.2e}  rel sigma(cpu) {np.max(np.sqrt(ov) / np.maximum(om, 1e-12)):.1e}"
          f"  (oracle {t1 - t0:.0f} s){flag}", flush=True)
sys.exit(1 if bad else 0)
