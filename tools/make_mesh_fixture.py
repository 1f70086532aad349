"""
Writes the small triangle meshes the canopy tests use (tests/golden/mesh_*.{ply,obj}) and asks the COMPILED REFERENCE
(oracle/_ref, see oracle/build_ref.sh) what its `ply` / `obj` plugins make of them: vertex / face buffers after
loading with a scaling `to_world` (what MeshTreeElement emits, _tree.py:470-478), i.e. the de-duplicated vertices,
the fan-triangulated faces and the angle-weighted vertex normals of mesh.cpp:330-382
-> tests/golden/mesh_reference.json.  Run here:  python tools/make_mesh_fixture.py
"""

import json
import os
import struct
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def icosphere(subdiv: int = 1):
    t = (1.0 + 5.0**0.5) / 2.0
    v = [(-1, t, 0), (1, t, 0), (-1, -t, 0), (1, -t, 0), (0, -1, t), (0, 1, t), (0, -1, -t), (0, 1, -t),
         (t, 0, -1), (t, 0, 1), (-t, 0, -1), (-t, 0, 1)]
    f = [(0, 11, 5), (0, 5, 1), (0, 1, 7), (0, 7, 10), (0, 10, 11), (1, 5, 9), (5, 11, 4), (11, 10, 2), (10, 7, 6),
         (7, 1, 8), (3, 9, 4), (3, 4, 2), (3, 2, 6), (3, 6, 8), (3, 8, 9), (4, 9, 5), (2, 4, 11), (6, 2, 10),
         (8, 6, 7), (9, 8, 1)]
    v = [np.array(x, dtype=np.float64) / np.linalg.norm(x) for x in v]
    for _ in range(subdiv):
        mid, nf = {}, []

        def m(a, b):
            key = (min(a, b), max(a, b))
            if key not in mid:
                p = v[a] + v[b]
                v.append(p / np.linalg.norm(p))
                mid[key] = len(v) - 1
            return mid[key]

        for a, b, c in f:
            ab, bc, ca = m(a, b), m(b, c), m(c, a)
            nf += [(a, ab, ca), (b, bc, ab), (c, ca, bc), (ab, bc, ca)]
        f = nf
    return np.array(v), np.array(f, dtype=np.uint32)


def crown():
    """An ellipsoidal crown (80 triangles), in CENTIMETRES (the tests load it with to_world = scale(0.01))."""
    v, f = icosphere(1)
    v = v * np.array([90.0, 70.0, 120.0]) + np.array([0.0, 0.0, 330.0])
    return v.astype(np.float32), f


def trunk():
    """A hexagonal prism, sides as quads and caps as hexagons (exercises the OBJ fan triangulation), centimetres."""
    ang = np.arange(6) * np.pi / 3
    ring = np.stack([12.0 * np.cos(ang), 12.0 * np.sin(ang)], axis=1)
    v = [(x, y, 0.0) for x, y in ring] + [(x, y, 215.0) for x, y in ring]
    faces = [[i + 1, (i + 1) % 6 + 1, (i + 1) % 6 + 7, i + 7] for i in range(6)]
    faces += [[6, 5, 4, 3, 2, 1], [7, 8, 9, 10, 11, 12]]
    return np.array(v, dtype=np.float32), faces


def leaf_quad():
    """A tilted quad of two triangles WITH vertex normals in the file (not the face normal: a curled leaf), metres."""
    v = np.array([(-0.2, -0.1, 0.5), (0.2, -0.1, 0.55), (0.2, 0.1, 0.65), (-0.2, 0.1, 0.6)], dtype=np.float32)
    n = np.array([(-0.3, -0.2, 1.0), (0.3, -0.2, 1.0), (0.3, 0.2, 1.0), (-0.3, 0.2, 1.0)], dtype=np.float32)
    n /= np.linalg.norm(n, axis=1)[:, None]
    return v, n, np.array([(0, 1, 2), (0, 2, 3)], dtype=np.uint32)


def write_ply(path, v, f, normals=None, fmt="binary_little_endian", extra_vertex_prop=False, extra_element=False):
    props = ["x", "y", "z"] + (["nx", "ny", "nz"] if normals is not None else []) + (["quality"] if extra_vertex_prop else [])
    head = ["ply", f"format {fmt} 1.0", "comment written by tools/make_mesh_fixture.py", f"element vertex {len(v)}"]
    head += [f"property float {p}" for p in props]
    if extra_element:
        head += ["element material 2", "property uchar red", "property uchar green"]
    head += [f"element face {len(f)}", "property list uchar int vertex_indices", "end_header"]
    rows = np.concatenate([v] + ([normals] if normals is not None else []) +
                          ([np.full((len(v), 1), 0.5, dtype=np.float32)] if extra_vertex_prop else []), axis=1)
    with open(path, "wb") as fh:
        fh.write(("\n".join(head) + "\n").encode())
        if fmt == "ascii":
            for r in rows:
                fh.write((" ".join(repr(float(x)) for x in r) + "\n").encode())
            if extra_element:
                fh.write(b"10 20\n30 40\n")
            for t in f:
                fh.write(("3 " + " ".join(str(int(i)) for i in t) + "\n").encode())
        else:
            e = "<" if fmt == "binary_little_endian" else ">"
            fh.write(rows.astype(e + "f4").tobytes())
            if extra_element:
                fh.write(bytes([10, 20, 30, 40]))
            for t in f:
                fh.write(struct.pack(e + "B3i", 3, *[int(i) for i in t]))


def write_obj(path, v, faces, normals=None):
    with open(path, "w") as fh:
        fh.write("# written by tools/make_mesh_fixture.py\n")
        for p in v:
            fh.write("v " + " ".join(repr(float(x)) for x in p) + "\n")
        if normals is not None:
            for n in normals:
                fh.write("vn " + " ".join(repr(float(x)) for x in n) + "\n")
        for face in faces:
            fh.write("f " + " ".join(f"{i}//{i}" if normals is not None else str(i) for i in face) + "\n")


def write_all(out_dir=GOLDEN):
    v, f = crown()
    write_ply(os.path.join(out_dir, "mesh_crown.ply"), v, f)
    v, faces = trunk()
    write_obj(os.path.join(out_dir, "mesh_trunk.obj"), v, faces)
    v, n, f = leaf_quad()
    write_ply(os.path.join(out_dir, "mesh_leaf_normals_ascii.ply"), v, f, normals=n, fmt="ascii", extra_vertex_prop=True)


CASES = [  # (file, plugin, to_world scale)
    ("mesh_crown.ply", "ply", 0.01),
    ("mesh_trunk.obj", "obj", 0.01),
    ("mesh_leaf_normals_ascii.ply", "ply", 1.0),
]


def main():
    write_all()
    from oracle import ref

    mi = ref.mitsuba("scalar_mono_double")
    out = {"generator": "tools/make_mesh_fixture.py", "reference": ref.describe(), "cases": []}
    for fname, plugin, scale in CASES:
        for face_normals in (False, True):
            sdict = {"type": plugin, "filename": os.path.join(GOLDEN, fname), "face_normals": face_normals,
                     "to_world": mi.ScalarTransform4f().scale(scale)}
            shape = mi.load_dict(sdict)
            scene = mi.load_dict({"type": "scene", "mesh": sdict})  # (a shape alone cannot be ray traced)
            p = mi.traverse(shape)
            case = {"file": fname, "plugin": plugin, "scale": scale, "face_normals": face_normals,
                    "vertex_count": int(shape.vertex_count()), "face_count": int(shape.face_count()),
                    "faces": np.array(p["faces"]).astype(int).tolist(),
                    "vertex_positions": np.array(p["vertex_positions"], dtype=np.float64).tolist(),
                    "vertex_normals": np.array(p["vertex_normals"], dtype=np.float64).tolist()}
            # and a few ray casts with the shading normal the plugin returns
            rng = np.random.default_rng(7)
            bb = shape.bbox()
            lo, hi = np.array(bb.min), np.array(bb.max)
            rays = []
            for _ in range(400):
                o = lo + (hi - lo) * rng.uniform(-0.3, 1.3, 3)
                tgt = lo + (hi - lo) * rng.uniform(0.1, 0.9, 3)
                d = tgt - o
                d /= np.linalg.norm(d)
                si = scene.ray_intersect(mi.Ray3f(mi.Point3f(*o), mi.Vector3f(*d)))
                if si.is_valid():
                    rays.append({"o": o.tolist(), "d": d.tolist(), "t": float(si.t),
                                 "n": [float(x) for x in si.n], "sh_n": [float(x) for x in si.sh_frame.n]})
            case["rays"] = rays[:60]
            out["cases"].append(case)
            print(fname, "face_normals" if face_normals else "smooth", case["vertex_count"], case["face_count"], len(rays))
    with open(os.path.join(GOLDEN, "mesh_reference.json"), "w") as fh:
        json.dump(out, fh)
    print("wrote mesh_reference.json", os.path.getsize(os.path.join(GOLDEN, "mesh_reference.json")), "bytes")


if __name__ == "__main__":
    main()
