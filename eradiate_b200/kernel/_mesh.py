"""
Triangle meshes of canopy elements (``MeshTreeElement``, src/eradiate/scenes/biosphere/_tree.py:285-470: a `ply` or
`obj` shape with a bilambertian BSDF and a scaling `to_world`): the two file readers and the vertex-normal
computation of the reference, on the host.  The device and the oracle receive plain triangles
(v0, v1, v2, n0, n1, n2).

* `ply`: MI/src/shapes/ply.cpp:155-435 + the header parser MI/src/render/mesh.cpp (ascii, binary little / big endian;
  vertex properties x y z [nx ny nz], anything else skipped; one list property `vertex_index` / `vertex_indices`
  per face, triangles only; unknown elements skipped; trailing content is an error);
* `obj`: MI/src/shapes/obj.cpp:150-408 (v / vn / f records; vertices de-duplicated by their (v, vt, vn) key;
  polygons fan-triangulated);
* vertex normals when the file has none and `face_normals` is false: angle-weighted face normals
  (Thuermer & Wuethrich 1998), MI/src/render/mesh.cpp:330-382, in float32 as the reference (`InputFloat`);
* shading: MI/src/render/mesh.cpp:1500-1560 -- interpolated vertex normals, the geometric normal with
  `face_normals`; `flip_normals` negates both.
"""

from __future__ import annotations

import os

import numpy as np

_PLY_TYPES = {
    "char": "i1", "int8": "i1", "uchar": "u1", "uint8": "u1", "short": "i2", "int16": "i2", "ushort": "u2", "uint16": "u2",
    "int": "i4", "int32": "i4", "uint": "u4", "uint32": "u4", "float": "f4", "float32": "f4", "double": "f8", "float64": "f8",
}


def _fail(kind: str, path: str, what: str):
    raise RuntimeError(f'Error while loading {kind} file "{os.path.basename(path)}": {what}!')


def read_ply(path: str):
    """-> (positions [n, 3] float32, normals [n, 3] float32 or None, faces [m, 3] uint32)."""
    fail = lambda what: _fail("PLY", path, what)  # noqa: E731
    if not os.path.exists(path):
        fail("file not found")
    raw = open(path, "rb").read()
    end = raw.find(b"end_header")
    if not raw.startswith(b"ply") or end < 0:
        fail("invalid PLY header")
    nl = raw.find(b"\n", end)
    header, body = raw[:end].decode("ascii", "replace").splitlines(), raw[nl + 1:]
    fmt, elements = None, []  # elements: [name, count, [(prop name, type) | (prop name, (count type, item type))]]
    for line in header[1:]:
        tok = line.split()
        if not tok or tok[0] in ("comment", "obj_info"):
            continue
        if tok[0] == "format":
            fmt = tok[1]
        elif tok[0] == "element":
            elements.append([tok[1], int(tok[2]), []])
        elif tok[0] == "property":
            if not elements:
                fail("property without an element")
            if tok[1] == "list":
                if tok[2] not in _PLY_TYPES or tok[3] not in _PLY_TYPES:
                    fail("unknown property type")
                elements[-1][2].append((tok[4], (_PLY_TYPES[tok[2]], _PLY_TYPES[tok[3]])))
            else:
                if tok[1] not in _PLY_TYPES:
                    fail("unknown property type")
                elements[-1][2].append((tok[2], _PLY_TYPES[tok[1]]))
        else:
            fail(f'invalid PLY header: unknown token "{tok[0]}"')
    if fmt not in ("ascii", "binary_little_endian", "binary_big_endian"):
        fail("invalid PLY header: unknown format")
    order = ">" if fmt == "binary_big_endian" else "<"
    positions = normals = faces = None
    pos = 0
    tokens = body.split() if fmt == "ascii" else None

    for name, count, props in elements:
        has_list = any(isinstance(t, tuple) for _, t in props)
        if name == "vertex":
            if has_list:
                fail("incompatible contents -- is this a triangle mesh?")
            names = [p for p, _ in props]
            if not all(k in names for k in ("x", "y", "z")):
                fail('Unable to find field "x"')
            if fmt == "ascii":
                n_tok = count * len(props)
                try:
                    arr = np.array(tokens[pos:pos + n_tok], dtype=np.float64).reshape(count, len(props))
                except ValueError:
                    fail("could not parse the vertex data")
                pos += n_tok
                col = {p: arr[:, i] for i, p in enumerate(names)}
            else:
                dt = np.dtype([(p, order + t) for p, t in props])
                if pos + count * dt.itemsize > len(body):
                    fail("read past the end of the file")
                rec = np.frombuffer(body, dtype=dt, count=count, offset=pos)
                pos += count * dt.itemsize
                col = {p: rec[p] for p in names}
            positions = np.stack([col["x"], col["y"], col["z"]], axis=1).astype(np.float32)
            if all(k in col for k in ("nx", "ny", "nz")):
                normals = np.stack([col["nx"], col["ny"], col["nz"]], axis=1).astype(np.float32)
        elif name == "face":
            lists = [(p, t) for p, t in props if isinstance(t, tuple)]
            if len(lists) != 1 or lists[0][0] not in ("vertex_index", "vertex_indices"):
                fail("vertex_index/vertex_indices property not found")
            if fmt == "ascii":
                width = len(props) + 3  # count + three indices, plus the scalar properties
                try:
                    arr = np.array(tokens[pos:pos + count * width], dtype=np.float64).reshape(count, width)
                except ValueError:
                    fail("incompatible contents -- is this a triangle mesh?")
                pos += count * width
                k = [p for p, _ in props].index(lists[0][0])
                if not np.all(arr[:, k] == 3):
                    fail("incompatible contents -- is this a triangle mesh?")
                faces = arr[:, k + 1:k + 4].astype(np.uint32)
            else:
                fields = []
                for p, t in props:
                    if isinstance(t, tuple):
                        fields += [("_n", order + t[0]), ("i", order + t[1], (3,))]
                    else:
                        fields.append((p, order + t))
                dt = np.dtype(fields)
                if pos + count * dt.itemsize > len(body):
                    fail("incompatible contents -- is this a triangle mesh?")
                rec = np.frombuffer(body, dtype=dt, count=count, offset=pos)
                pos += count * dt.itemsize
                if not np.all(rec["_n"] == 3):
                    fail("incompatible contents -- is this a triangle mesh?")
                faces = rec["i"].astype(np.uint32)
        else:  # unknown element: skipped (fixed-size records only, as the reference's seek)
            if has_list:
                fail(f'cannot skip element "{name}" (list property)')
            if fmt == "ascii":
                pos += count * len(props)
            else:
                pos += count * np.dtype([(p, order + t) for p, t in props]).itemsize
    if fmt == "ascii":
        if pos != len(tokens):
            fail("invalid file -- trailing content")
    elif pos != len(body):
        fail("invalid file -- trailing content")
    if positions is None or faces is None:
        fail("vertex or face element not found")
    if faces.size and faces.max() >= positions.shape[0]:
        fail("face refers to an invalid vertex")
    return positions, normals, faces


def read_obj(path: str):
    """-> (positions, normals or None, faces) with the reference's de-duplication of (v, vt, vn) keys."""
    fail = lambda what: _fail("OBJ", path, what)  # noqa: E731
    if not os.path.exists(path):
        fail("file not found")
    v, vn, keys, key_id, tris = [], [], [], {}, []
    for line in open(path, "r", errors="replace"):
        tok = line.split()
        if not tok:
            continue
        try:
            if tok[0] == "v":
                v.append([float(x) for x in tok[1:4]])
                if len(v[-1]) != 3:
                    raise ValueError
            elif tok[0] == "vn":
                vn.append([float(x) for x in tok[1:4]])
                if len(vn[-1]) != 3:
                    raise ValueError
            elif tok[0] == "f":
                ids = []
                for item in tok[1:]:
                    parts = item.split("/")
                    if len(parts) > 3:
                        raise ValueError
                    key = tuple(int(x) if x else 0 for x in parts) + (0,) * (3 - len(parts))
                    if key[0] < 1 or key[0] > len(v):
                        fail(f"reference to invalid vertex {key[0]}")
                    if key not in key_id:
                        key_id[key] = len(keys)
                        keys.append(key)
                    ids.append(key_id[key])
                for k in range(2, len(ids)):  # obj.cpp:312-320: a fan around the first vertex
                    tris.append([ids[0], ids[k - 1], ids[k]])
        except ValueError:
            fail(f'could not parse line "{line.strip()}"')
    positions = np.array([v[k[0] - 1] for k in keys], dtype=np.float32).reshape(-1, 3)
    normals = None
    if vn:
        if any(k[2] == 0 for k in keys):
            fail("vertices with and without normals in one file are not supported")
        if any(k[2] > len(vn) for k in keys):
            fail("reference to invalid normal")
        normals = np.array([vn[k[2] - 1] for k in keys], dtype=np.float32).reshape(-1, 3)
    return positions, normals, np.array(tris, dtype=np.uint32).reshape(-1, 3)


def _unit_angle(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """dr::unit_angle: numerically robust angle between unit vectors (float32)."""
    dot = np.sum(a * b, axis=1)
    t = 2.0 * np.arcsin(np.minimum(0.5 * np.linalg.norm(b - np.where((dot < 0)[:, None], -a, a), axis=1), 1.0).astype(np.float32))
    return np.where(dot >= 0, t, np.float32(np.pi) - t).astype(np.float32)


def vertex_normals(positions: np.ndarray, faces: np.ndarray) -> np.ndarray:
    """mesh.cpp:330-382 (scalar branch), float32."""
    p = positions.astype(np.float32)
    v0, v1, v2 = p[faces[:, 0]], p[faces[:, 1]], p[faces[:, 2]]
    n = np.cross(v1 - v0, v2 - v0).astype(np.float32)
    l2 = np.sum(n * n, axis=1)
    ok = l2 > 0
    n[ok] /= np.sqrt(l2[ok])[:, None]

    def unit(x):
        return (x / np.maximum(np.linalg.norm(x, axis=1), np.float32(1e-38))[:, None]).astype(np.float32)

    ang = np.stack([_unit_angle(unit(v1 - v0), unit(v2 - v0)), _unit_angle(unit(v2 - v1), unit(v0 - v1)),
                    _unit_angle(unit(v0 - v2), unit(v1 - v2))], axis=1)
    out = np.zeros_like(p)
    for j in range(3):
        np.add.at(out, faces[ok, j], n[ok] * ang[ok, j][:, None])
    ln = np.linalg.norm(out, axis=1)
    bad = ln == 0
    out[~bad] /= ln[~bad][:, None]
    out[bad] = (1.0, 0.0, 0.0)  # "some bogus value", :372
    return out.astype(np.float32)


def load_triangles(kind: str, path: str, to_world: np.ndarray, face_normals: bool = False, flip_normals: bool = False):
    """[m, 18] float32 rows (v0, v1, v2, n0, n1, n2) in the shape's world space, plus (vertex_count, face_count)."""
    positions, normals, faces = (read_ply if kind == "ply" else read_obj)(path)
    m = np.asarray(to_world, dtype=np.float64)
    hp = positions.astype(np.float64) @ m[:3, :3].T + m[:3, 3]
    w = positions.astype(np.float64) @ m[3, :3] + m[3, 3]
    positions = (hp / w[:, None]).astype(np.float32)
    if not np.all(np.isfinite(positions)):
        _fail(kind.upper(), path, "mesh contains invalid vertex position data")
    if face_normals:
        normals = None
    elif normals is not None:  # Transform * Normal: the inverse transpose
        nt = normals.astype(np.float64) @ np.linalg.inv(m[:3, :3])
        normals = (nt / np.maximum(np.linalg.norm(nt, axis=1), 1e-300)[:, None]).astype(np.float32)
    else:
        normals = vertex_normals(positions, faces)
    v = [positions[faces[:, k]] for k in range(3)]
    if normals is None:
        g = np.cross(v[1] - v[0], v[2] - v[0]).astype(np.float64)
        g /= np.maximum(np.linalg.norm(g, axis=1), 1e-300)[:, None]
        nn = [g.astype(np.float32)] * 3
    else:
        nn = [normals[faces[:, k]] for k in range(3)]
    rows = np.concatenate(v + nn, axis=1).astype(np.float32)
    if flip_normals:
        rows[:, 9:] *= -1.0
    # degenerate triangles can never be hit (Moeller-Trumbore: 1 / 0 determinant): dropped here
    area2 = np.linalg.norm(np.cross((v[1] - v[0]).astype(np.float64), (v[2] - v[0]).astype(np.float64)), axis=1)
    return np.ascontiguousarray(rows[area2 > 0]), (positions.shape[0], faces.shape[0])
