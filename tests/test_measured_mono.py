"""
`measured_mono` (ERP/bsdfs/measured_mono.cpp): every CPU-side piece against values the COMPILED REFERENCE returned
for two synthetic RGL tensor files (tests/golden/measured_{iso,aniso}.bsdf, written by tools/make_measured_fixture.py
together with tests/golden/measured_mono_reference.json: BSDF::eval / pdf / sample of the `scalar_mono_double`
variant at 48 direction pairs, two wavelengths per file).

* oracle/measured_mono.py: numpy restatement working from the tensor file (its own table construction);
* eradiate_b200/kernel/_measured.py: the product's loader, which flattens the file into one float32 table;
* oracle/ertb_oracle_measured.c: the C oracle, which evaluates that table in double precision.
The C oracle reading the product's table and matching the reference therefore pins the flattening as well.
The reference's own test of the plugin (ERP/tests/bsdfs/test_measured_mono.py) needs a material of the RGL database
that is not in the tree; the chi-square property it checks (sample() follows pdf()) is tested here on the synthetic
files instead.
"""

import json
import os

import numpy as np
import pytest

from eradiate_b200 import scenes
from eradiate_b200.kernel import mi_load_dict, mi_traverse
from eradiate_b200.kernel._measured import HEADER, MeasuredData, read_tensor_file

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
REF = json.load(open(os.path.join(GOLDEN, "measured_mono_reference.json")))
CASES = [(c["file"], c["wavelength"]) for c in REF["cases"]]


def surface(fname, wavelength):
    return {"type": "measured_mono", "filename": os.path.join(GOLDEN, fname), "wavelength": wavelength}


def make(fname, wavelength, **kw):
    kw.setdefault("geometry", "plane_parallel")
    kw.setdefault("n_layers", 10)
    sc = mi_load_dict(scenes.atmosphere_scene(surface=surface(fname, wavelength), **kw))
    return sc, sc.flat.build_desc()


def case(fname, wavelength):
    return next(c for c in REF["cases"] if c["file"] == fname and c["wavelength"] == wavelength)


@pytest.mark.parametrize("fname,wavelength", CASES)
def test_numpy_restatement_matches_the_reference(fname, wavelength):
    from oracle.measured_mono import MeasuredMono

    c = case(fname, wavelength)
    b = MeasuredMono(os.path.join(GOLDEN, fname), wavelength)
    for k, (wi, wo, u) in enumerate(zip(c["wi"], c["wo"], c["u"])):
        assert np.isclose(b.eval(wi, wo), c["eval"][k], rtol=2e-6, atol=1e-9)
        assert np.isclose(b.pdf(wi, wo), c["pdf"][k], rtol=2e-6, atol=1e-9)
        swo, w, pdf = b.sample(wi, u)
        assert np.allclose(swo, c["sample_wo"][k], atol=5e-6)
        assert np.isclose(w, c["sample_weight"][k], rtol=2e-5, atol=1e-9)
        assert np.isclose(pdf, c["sample_pdf"][k], rtol=2e-5, atol=1e-9)


@pytest.mark.parametrize("fname,wavelength", CASES)
def test_c_oracle_on_the_host_table_matches_the_reference(oracle, fname, wavelength):
    c = case(fname, wavelength)
    _, desc = make(fname, wavelength)
    wi, wo, u = np.array(c["wi"]), np.array(c["wo"]), np.array(c["u"])
    # float32 table (the reference holds float32 data too, but accumulates its CDFs in double): 2e-6 relative
    assert np.allclose(oracle.bsdf_eval(desc, wi, wo), c["eval"], rtol=5e-6, atol=1e-9)
    u3 = np.concatenate([np.full((len(u), 1), 0.5), u], axis=1)  # (sample1, sample2.x, sample2.y)
    swo, w = oracle.bsdf_sample(desc, wi, u3)
    assert np.allclose(swo, c["sample_wo"], atol=2e-5)
    assert np.allclose(w, c["sample_weight"], rtol=5e-5, atol=1e-9)


def test_sampling_follows_the_pdf(oracle):
    """The property ERP/tests/bsdfs/test_measured_mono.py checks with a chi-square test -- sample() is distributed
    as pdf() -- through importance-weighted integrals: E[g(wo) / pdf(wo)] = int g over the upper hemisphere, for a
    few test functions g (pdf() has an integrable singularity at the mirror direction, so a histogram test on a
    direction grid would need a very fine quadrature there)."""
    for fname, wavelength in CASES[::2]:
        c = case(fname, wavelength)
        _, desc = make(fname, wavelength)
        # the C pdf against the reference's
        assert np.allclose(oracle.bsdf_pdf(desc, np.array(c["wi"]), np.array(c["wo"])), c["pdf"], rtol=5e-6, atol=1e-9)
        wi = np.array([np.sin(0.6) * np.cos(0.4), np.sin(0.6) * np.sin(0.4), np.cos(0.6)])
        n = 200000
        u = np.random.default_rng(5).uniform(0, 1, (n, 3))
        wis = np.tile(wi, (n, 1))
        wo, w = oracle.bsdf_sample(desc, wis, u)
        pdf = oracle.bsdf_pdf(desc, wis, wo)
        up = wo[:, 2] > 0
        assert np.all(pdf[up] > 0)
        for g, exact in ((lambda d: np.ones(len(d)), 2 * np.pi), (lambda d: d[:, 2], np.pi),
                         (lambda d: d[:, 0], 0.0), (lambda d: d[:, 1] ** 2, 2 * np.pi / 3)):
            x = np.where(up, g(wo) / np.where(up, pdf, 1.0), 0.0)
            est, err = x.mean(), x.std() / np.sqrt(n)
            assert abs(est - exact) < 5 * err + 1e-3, (fname, est, exact, err)
        # and the weight is eval / pdf
        ev = oracle.bsdf_eval(desc, wis[up][:2000], wo[up][:2000])
        assert np.allclose(w[up][:2000], ev / pdf[up][:2000], rtol=1e-5)  # (invert(sample(s)) = s up to rounding)


def test_table_layout_and_wavelength_blend():
    path = os.path.join(GOLDEN, "measured_aniso.bsdf")
    tf = read_tensor_file(path)
    md = MeasuredData(path)
    assert (md.isotropic, md.reduction, md.jacobian) == (False, 2, int(tf["jacobian"][0]))
    wv = tf["wavelengths"]
    # at a node the blend is that slice; between nodes it is linear; outside the range it clamps (distr_2d.h:280-284)
    n_phi, n_theta = tf["phi_i"].size, tf["theta_i"].size
    flat = lambda a: a.reshape((n_phi * n_theta,) + a.shape[2:])  # noqa: E731
    assert np.allclose(md.spectra_at(float(wv[1])), flat(tf["spectra"][:, :, 1]))
    mid = 0.25 * wv[1] + 0.75 * wv[2]
    assert np.allclose(md.spectra_at(float(mid)), flat(0.25 * tf["spectra"][:, :, 1] + 0.75 * tf["spectra"][:, :, 2]), rtol=1e-6)
    assert np.allclose(md.spectra_at(float(wv[0]) - 50.0), flat(tf["spectra"][:, :, 0]))
    assert np.allclose(md.spectra_at(float(wv[-1]) + 50.0), flat(tf["spectra"][:, :, -1]))
    T = md.table(600.0)
    assert T.dtype == np.float32 and T.size > HEADER
    h = T[:HEADER].astype(int)
    assert (h[0], h[1], h[2], h[4]) == (n_phi, n_theta, 0, 2) and (h[26], h[27]) == (n_theta, 1)
    H, W = tf["vndf"].shape[2:]
    assert (h[11], h[12], h[16], h[17], h[21], h[22]) == (W, H, W, H, W, H)
    # normalised interpolants: the marginal CDF of every slice ends at 1
    marg = T[h[14]:h[14] + n_phi * n_theta * (H - 1)].reshape(-1, H - 1)
    assert np.allclose(marg[:, -1], 1.0, atol=1e-6)
    assert np.array_equal(T[h[24]:h[24] + n_phi], tf["phi_i"]) and np.array_equal(T[h[25]:h[25] + n_theta], tf["theta_i"])
    assert h[25] + n_theta == T.size


def test_loader_errors_and_traversal(tmp_path):
    with pytest.raises(RuntimeError, match="filename"):
        mi_load_dict(scenes.atmosphere_scene(surface={"type": "measured_mono"}, n_layers=4))
    bad = tmp_path / "bad.bsdf"
    bad.write_bytes(b"not a tensor file at all")
    with pytest.raises(RuntimeError, match="Invalid tensor file"):
        mi_load_dict(scenes.atmosphere_scene(surface={"type": "measured_mono", "filename": str(bad)}, n_layers=4))
    # an RGB measurement (no "wavelengths" field): measured_mono.cpp:72-83
    import tools.make_measured_fixture as mk

    fields = {k: v for k, v in mk.synthetic(1, seed=3).items() if k != "wavelengths"}
    rgb = tmp_path / "rgb.bsdf"
    mk.write_tensor_file(str(rgb), fields)
    with pytest.raises(RuntimeError, match="RGB format"):
        mi_load_dict(scenes.atmosphere_scene(surface={"type": "measured_mono", "filename": str(rgb)}, n_layers=4))
    sc, _ = make("measured_iso.bsdf", 450.0)
    params = mi_traverse(sc).parameters
    assert params["surface_bsdf.wavelength"] == 450.0  # the only parameter the plugin publishes (:224-226)
    assert [k for k in params.keys() if k.startswith("surface_bsdf.")] == ["surface_bsdf.wavelength"]


def test_wavelength_update_rebuilds_the_table():
    sc, d0 = make("measured_iso.bsdf", 450.0)
    t0 = np.ctypeslib.as_array(d0.bsdf_table, (d0.bsdf_table_res[0],)).copy()
    mi_traverse(sc).parameters.update({"surface_bsdf.wavelength": 725.0})
    d1 = sc.flat.build_desc()
    t1 = np.ctypeslib.as_array(d1.bsdf_table, (d1.bsdf_table_res[0],))
    ref = MeasuredData(os.path.join(GOLDEN, "measured_iso.bsdf")).table(725.0)
    assert np.array_equal(t1, ref) and not np.array_equal(t0, t1)
    assert np.array_equal(t0[:HEADER], t1[:HEADER])  # only the blended spectra change
