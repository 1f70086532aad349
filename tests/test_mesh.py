"""
Mesh canopy elements (`ply` / `obj` shapes of MeshTreeElement, src/eradiate/scenes/biosphere/_tree.py:285-478) on
the CPU side: the file readers and the vertex normals of eradiate_b200/kernel/_mesh.py, and the triangle ray caster
of the oracle, against what the COMPILED REFERENCE made of the same files (tests/golden/mesh_reference.json, written
by tools/make_mesh_fixture.py: vertex / face buffers of MI/src/shapes/ply.cpp and obj.cpp after loading, ray casts
with geometric and shading normals from MI/src/render/mesh.cpp).
"""

import json
import os

import numpy as np
import pytest

from eradiate_b200 import scenes
from eradiate_b200.kernel import mi_load_dict, mi_traverse
from eradiate_b200.kernel._mesh import load_triangles, read_obj, read_ply, vertex_normals

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
REF = json.load(open(os.path.join(GOLDEN, "mesh_reference.json")))
IDS = [f"{c['file']}-{'faceted' if c['face_normals'] else 'smooth'}" for c in REF["cases"]]


def mesh_scene(elements, positions=((0.0, 0.0),), **kw):
    kw.setdefault("sensor", {"type": "mdistant", "vza": [0.0], "vaa": 0.0})
    return mi_load_dict(scenes.atmosphere_scene(
        geometry="plane_parallel", atmosphere=None, integrator="path",
        canopy={"mesh_trees": {"elements": elements, "positions": positions}, "size": (4.0, 4.0, 4.6)}, **kw))


def element(c, **kw):
    return dict({"id": "m", "filename": os.path.join(GOLDEN, c["file"]), "scale": c["scale"], "reflectance": 0.5,
                 "transmittance": 0.3, "face_normals": c["face_normals"]}, **kw)


@pytest.mark.parametrize("c", REF["cases"], ids=IDS)
def test_buffers_match_the_reference(c):
    path = os.path.join(GOLDEN, c["file"])
    pos, nrm, faces = (read_ply if c["plugin"] == "ply" else read_obj)(path)
    ref_pos = np.array(c["vertex_positions"]).reshape(-1, 3)
    assert (pos.shape[0], faces.shape[0]) == (c["vertex_count"], c["face_count"])
    assert np.array_equal(faces, np.array(c["faces"]).reshape(-1, 3))  # incl. the OBJ fan triangulation / de-duplication
    assert np.allclose(pos * c["scale"], ref_pos, rtol=3e-7, atol=1e-9)
    rows, counts = load_triangles(c["plugin"], path, np.diag([c["scale"]] * 3 + [1.0]), face_normals=c["face_normals"])
    assert counts == (c["vertex_count"], c["face_count"]) and rows.shape == (c["face_count"], 18)
    for k in range(3):
        assert np.allclose(rows[:, 3 * k:3 * k + 3], ref_pos[faces[:, k]], rtol=3e-7, atol=1e-9)
    if c["face_normals"]:
        assert not c["vertex_normals"]  # the reference keeps no vertex normals then
        g = np.cross(rows[:, 3:6] - rows[:, 0:3], rows[:, 6:9] - rows[:, 0:3])
        g /= np.linalg.norm(g, axis=1)[:, None]
        for k in range(3):
            assert np.allclose(rows[:, 9 + 3 * k:12 + 3 * k], g, atol=1e-6)
    else:
        ref_n = np.array(c["vertex_normals"]).reshape(-1, 3)
        for k in range(3):  # file normals, or the angle-weighted ones of mesh.cpp:330-382
            assert np.allclose(rows[:, 9 + 3 * k:12 + 3 * k], ref_n[faces[:, k]], atol=5e-7)


@pytest.mark.parametrize("c", REF["cases"], ids=IDS)
def test_oracle_ray_caster_matches_the_reference(oracle, c):
    desc = mesh_scene([element(c)]).flat.build_desc()
    o = np.array([r["o"] for r in c["rays"]])
    d = np.array([r["d"] for r in c["rays"]])
    t, n, _ = oracle.canopy_intersect(desc, o, d)
    scale = max(1.0, np.abs(o).max())
    assert np.allclose(t, [r["t"] for r in c["rays"]], rtol=2e-6, atol=2e-6 * scale)  # (float32 vertices either side)
    assert np.allclose(n, [r["sh_n"] for r in c["rays"]], atol=2e-5)  # the shading normal (mesh.cpp:1500-1535)
    if not c["face_normals"] and c["file"] != "mesh_trunk.obj":
        assert not np.allclose([r["n"] for r in c["rays"]], [r["sh_n"] for r in c["rays"]], atol=1e-3)  # really interpolated


def test_ply_variants_and_errors(tmp_path):
    import tools.make_mesh_fixture as mk

    v, f = mk.crown()
    files = {}
    for fmt in ("ascii", "binary_little_endian", "binary_big_endian"):
        files[fmt] = str(tmp_path / f"crown_{fmt}.ply")
        mk.write_ply(files[fmt], v, f, fmt=fmt, extra_vertex_prop=(fmt != "ascii"), extra_element=True)
    ref = read_ply(os.path.join(GOLDEN, "mesh_crown.ply"))
    for fmt, path in files.items():  # unknown elements and vertex properties are skipped (ply.cpp:409-412)
        pos, nrm, faces = read_ply(path)
        assert nrm is None and np.array_equal(faces, ref[2]) and np.allclose(pos, ref[0], rtol=1e-7)
    # vertex normals in an OBJ
    lv, ln, lf = mk.leaf_quad()
    p = str(tmp_path / "leaf.obj")
    mk.write_obj(p, lv, [[int(i) + 1 for i in t] for t in lf], normals=ln)
    pos, nrm, faces = read_obj(p)
    assert np.allclose(pos, lv) and np.allclose(nrm, ln, atol=1e-7) and np.array_equal(faces, lf)
    # errors carry the reference's wording
    with pytest.raises(RuntimeError, match="file not found"):
        read_ply(str(tmp_path / "nope.ply"))
    quad = str(tmp_path / "quad.ply")
    open(quad, "w").write("ply\nformat ascii 1.0\nelement vertex 4\nproperty float x\nproperty float y\nproperty float z\n"
                          "element face 1\nproperty list uchar int vertex_indices\nend_header\n"
                          "0 0 0\n1 0 0\n1 1 0\n0 1 0\n4 0 1 2 3\n")
    with pytest.raises(RuntimeError, match="is this a triangle mesh"):
        read_ply(quad)
    trailing = str(tmp_path / "trailing.ply")
    open(trailing, "wb").write(open(files["binary_little_endian"], "rb").read() + b"\0\0")
    with pytest.raises(RuntimeError, match="trailing content"):
        read_ply(trailing)
    bad = str(tmp_path / "bad.obj")
    open(bad, "w").write("v 0 0 0\nv 1 0 0\nv 0 1 0\nf 1 2 7\n")
    with pytest.raises(RuntimeError, match="reference to invalid vertex 7"):
        read_obj(bad)
    # a degenerate face contributes nothing to the vertex normals (mesh.cpp:355) and is never hit
    vv = np.array([(0, 0, 0), (1, 0, 0), (0, 1, 0), (2, 0, 0)], dtype=np.float32)
    n = vertex_normals(vv, np.array([(0, 1, 2), (0, 1, 3)], dtype=np.uint32))
    assert np.allclose(n[:3], (0, 0, 1)) and np.allclose(n[3], (1, 0, 0))  # ("some bogus value", :372)


def test_loader_scene_graph_and_updates():
    c = REF["cases"][0]
    with pytest.raises(RuntimeError, match="inside a shapegroup only"):
        mi_load_dict(scenes.atmosphere_scene(geometry="plane_parallel", atmosphere=None, integrator="path")
                     | {"m": {"type": "ply", "filename": os.path.join(GOLDEN, c["file"])}})
    with pytest.raises(RuntimeError, match="bilambertian"):
        d = scenes.atmosphere_scene(geometry="plane_parallel", atmosphere=None, integrator="path",
                                    canopy={"mesh_trees": {"elements": [element(c)]}, "size": (4.0, 4.0, 4.6)})
        d["bsdf_m"] = {"type": "diffuse"}
        mi_load_dict(d)
    sc = mesh_scene([element(c, id="crown"), element(REF["cases"][2], id="trunk", reflectance=0.2, transmittance=0.0)],
                    positions=((0.0, 0.0), (2.0, 1.0)))
    d = sc.flat.build_desc()
    g = d.leaf_groups[0]
    assert (d.n_leaf_groups, d.n_instances, g.n_disks, g.n_triangles, g.n_mesh_bsdfs) == (1, 2, 0, 100, 2)
    assert np.allclose(np.ctypeslib.as_array(g.mesh_bsdfs, (2, 2)), [[0.5, 0.3], [0.2, 0.0]])
    ids = np.ctypeslib.as_array(g.triangle_bsdf, (100,))
    assert np.all(ids[:80] == 0) and np.all(ids[80:] == 1)
    w = mi_traverse(sc)
    keys = [k for k in w.parameters.keys() if k.startswith("bsdf_")]
    assert keys == ["bsdf_crown.reflectance.value", "bsdf_crown.transmittance.value",
                    "bsdf_trunk.reflectance.value", "bsdf_trunk.transmittance.value"]
    w.parameters.update({"bsdf_trunk.reflectance.value": 0.35})
    assert np.allclose(sc.flat.mesh_bsdf_params(0), [[0.5, 0.3], [0.35, 0.0]])


def test_flat_mesh_square_equals_its_closed_form(oracle):
    """A horizontal unit square of two triangles (face normals) over a black ground, no atmosphere, sun at the
    zenith, seen from the zenith by a sensor aimed at its centre: L = r E / pi (bilambertian reflection lobe)."""
    import tools.make_mesh_fixture as mk
    import tempfile

    with tempfile.TemporaryDirectory() as tmp:
        p = os.path.join(tmp, "square.ply")
        v = np.array([(-0.5, -0.5, 1.0), (0.5, -0.5, 1.0), (0.5, 0.5, 1.0), (-0.5, 0.5, 1.0)], dtype=np.float32)
        mk.write_ply(p, v, np.array([(0, 1, 2), (0, 2, 3)], dtype=np.uint32), fmt="ascii")
        sc = mi_load_dict(scenes.atmosphere_scene(
            geometry="plane_parallel", atmosphere=None, integrator="path", sza=0.0,
            surface={"type": "diffuse", "reflectance": 0.0},
            canopy={"mesh_trees": {"elements": [{"id": "sq", "filename": p, "reflectance": 0.4, "transmittance": 0.35,
                                                 "face_normals": True}], "positions": ((0.0, 0.0),)},
                    "size": (0.5, 0.5, 1.0)},
            sensor={"type": "mdistant", "vza": [0.0, 30.0], "vaa": 0.0}))
        desc = sc.flat.build_desc()
        spp = 1 << 12
        _, l, _, _ = oracle.render(desc, 0, 1, spp)
    e = desc.irradiance
    assert np.allclose(l / spp, 0.4 * e / np.pi, rtol=1e-9)  # every ray hits the square; nothing else scatters
