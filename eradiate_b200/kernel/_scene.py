"""
Host-side scene graph and flattener.

``load_dict`` walks the very same nested plugin dictionary Eradiate passes to
``mitsuba.load_dict`` (``src/eradiate/kernel/_render.py:186-209``; the dict layout
is restated in SURVEY.md section 3.5) and builds a small node tree that

* exposes the parameter paths ``mi_traverse`` publishes (same dotted keys as
  Mitsuba's ``traverse`` callbacks: ``<shape>.interior_medium.sigma_t.volume.data``,
  ``<emitter>.irradiance.value``, ``<bsdf>.rho_0.value``, ``...phase_0.values`` ...),
* flattens to the POD ``ertb_scene_desc`` consumed by the CUDA library
  (``include/eradiate_b200.h``).

Unsupported plugins raise ``RuntimeError`` -- the convention of
``src/eradiate/experiments/_core.py:670-671``.  There is no CPU fallback.
"""

from __future__ import annotations

import ctypes as C
import typing as t

import numpy as np

from .. import _abi
from ._types import ScalarTransform4f, to_grid, to_matrix

# ------------------------------------------------------------------------------
#                               dict utilities
# ------------------------------------------------------------------------------


def unflatten(d: dict) -> dict:
    """Expand dotted keys (``"film.width": 32``) into nested dicts, recursively."""
    out: dict = {}
    for k, v in d.items():
        if isinstance(v, dict):
            v = unflatten(v)
        if isinstance(k, str) and "." in k and not _looks_like_id(k, v):
            head, rest = k.split(".", 1)
            node = out.setdefault(head, {})
            if not isinstance(node, dict):
                raise RuntimeError(f"conflicting dictionary entries for '{head}'")
            _merge(node, unflatten({rest: v}))
        else:
            if k in out and isinstance(out[k], dict) and isinstance(v, dict):
                _merge(out[k], v)
            else:
                out[k] = v
    return out


def _looks_like_id(key: str, value) -> bool:
    # Top-level object ids never contain dots in Eradiate-emitted dicts.
    return False


def _merge(dst: dict, src: dict) -> None:
    for k, v in src.items():
        if k in dst and isinstance(dst[k], dict) and isinstance(v, dict):
            _merge(dst[k], v)
        else:
            dst[k] = v


def _parse_floats(value, what: str) -> np.ndarray:
    """Comma/space separated float string (tabphase.cpp:45-62, mdistant.cpp:104-111)."""
    if isinstance(value, str):
        toks = value.replace(",", " ").split()
        try:
            return np.array([float(s) for s in toks], dtype=np.float64)
        except ValueError as e:
            raise RuntimeError(f"could not parse floating point value in '{what}': {e}") from e
    return np.asarray(value, dtype=np.float64).ravel()


def _scalar_spectrum(value, what: str) -> float:
    """A mono-mode spectrum/texture: float or ``{"type": "uniform", "value": x}``."""
    if isinstance(value, dict):
        ty = value.get("type")
        if ty != "uniform":
            raise RuntimeError(f"unsupported spectrum type '{ty}' for '{what}' (only 'uniform')")
        return float(value["value"])
    return float(value)


# ------------------------------------------------------------------------------
#                                 node classes
# ------------------------------------------------------------------------------


class Object:
    """Base scene-graph node (``mitsuba.Object`` stand-in)."""

    plugin_kind = "object"

    def __init__(self, type_: str, id_: str | None = None):
        self.type = type_
        self._id = id_ or ""
        self.children: dict[str, Object] = {}
        self.values: dict[str, t.Any] = {}
        self.dirty = True
        self.dirty_names: set[str] = set()  # values changed since the device scene last synchronised

    def id(self) -> str:
        return self._id

    def traverse(self, cb) -> None:
        for name, value in self.values.items():
            cb.put(name, value, 0)
        for name, child in self.children.items():
            cb.put(name, child, 0)

    def set_value(self, name: str, value) -> None:
        cur = self.values[name]
        if isinstance(cur, np.ndarray):
            new = np.asarray(value, dtype=cur.dtype)
            if new.size != cur.size:
                raise RuntimeError(
                    f"parameter '{name}' of {self}: size mismatch ({new.size} != {cur.size})"
                )
            self.values[name] = new.reshape(cur.shape).copy()
        else:
            self.values[name] = type(cur)(np.asarray(value).reshape(-1)[0])
        self.dirty = True
        self.dirty_names.add(name)

    def __repr__(self):
        return f"{type(self).__name__}[type={self.type}, id={self._id!r}]"


class Texture(Object):
    plugin_kind = "texture"


class Volume(Object):
    plugin_kind = "volume"

    def layer_values(self) -> np.ndarray:
        """1D profile carried by the volume (float32)."""
        if self.type == "constvolume":  # volumes/const.cpp:89: a texture child named `value`
            return np.array([self.children["value"].values["value"]], dtype=np.float32)
        if self.type == "sphericalcoordsvolume":
            return self.children["volume"].layer_values()
        data = self.values["data"]
        return data.reshape(-1).astype(np.float32)


class PhaseFunction(Object):
    plugin_kind = "phase"


class BSDF(Object):
    plugin_kind = "bsdf"


class Medium(Object):
    plugin_kind = "medium"


class Emitter(Object):
    plugin_kind = "emitter"


class Shape(Object):
    plugin_kind = "shape"


class Film(Object):
    plugin_kind = "film"

    def __init__(self, type_, id_=None):
        super().__init__(type_, id_)
        self._bitmap = None

    def size(self):
        return (self.width, self.height)

    def bitmap(self, raw: bool = False):
        if self._bitmap is None:
            raise RuntimeError("film has not been rendered yet")
        return self._bitmap


class Sampler(Object):
    plugin_kind = "sampler"


class Sensor(Object):
    plugin_kind = "sensor"

    def film(self) -> Film:
        return self.children["film"]

    def sampler(self) -> Sampler:
        return self.children["sampler"]


class Integrator(Object):
    plugin_kind = "integrator"


class Scene(Object):
    plugin_kind = "scene"

    def __init__(self):
        super().__init__("scene", "")
        self._sensors: list[Sensor] = []
        self._flat: FlatScene | None = None

    def sensors(self) -> list[Sensor]:
        return self._sensors

    def integrator(self) -> Integrator:
        return self.children[self._integrator_key]

    @property
    def flat(self) -> "FlatScene":
        if self._flat is None:
            self._flat = FlatScene(self)
        return self._flat


# ------------------------------------------------------------------------------
#                                   loader
# ------------------------------------------------------------------------------

_SUPPORTED = {
    "integrator": {"volpath", "volpathmis", "piecewise_volpath", "path", "moment", "stokes"},
    "emitter": {"directional", "astroobject"},
    "shape": {"sphere", "cube", "rectangle", "arectangle", "disk", "shapegroup", "instance", "cylinder", "ply", "obj"},
    "medium": {"heterogeneous", "homogeneous", "piecewise"},
    "bsdf": {"diffuse", "rpv", "rtls", "hapke", "ocean_legacy", "ocean_mishchenko", "ocean_grasp", "maignan",
             "mqdiffuse", "measured_mono", "null", "bilambertian", "blendbsdf", "selectbsdf"},
    "phase": {"isotropic", "rayleigh", "hg", "tabphase", "tabphase_irregular", "blendphase",
              "rayleigh_polarized", "tabphase_polarized", "multiphase"},
    "sensor": {"mdistant", "hdistant", "distantflux", "perspective", "mpdistant", "mradiancemeter"},
    "volume": {"gridvolume", "sphericalcoordsvolume", "constvolume"},
}
_KIND_OF = {ty: kind for kind, types in _SUPPORTED.items() for ty in types}
# plugins the reference ships for this slot but that this kernel does not (yet) implement
_KNOWN_UNSUPPORTED = {
}


def _phase_leaf_nodes(ph):
    """Leaves below a blendphase / multiphase node."""
    kids = [c for name, c in ph.children.items() if name.startswith("phase")]
    if ph.type not in ("blendphase", "multiphase"):
        return [ph]
    return [leaf for c in kids for leaf in _phase_leaf_nodes(c)]


class _Loader:
    def __init__(self, root: dict):
        self.root = unflatten(root)
        self.by_id: dict[str, Object] = {}
        self.dict_by_id: dict[str, dict] = {}
        self._index_ids(self.root, top=True)

    def _index_ids(self, d: dict, top: bool = False) -> None:
        for k, v in d.items():
            if isinstance(v, dict) and "type" in v and v["type"] != "ref":
                oid = v.get("id", k if top else None)
                if oid is not None:
                    self.dict_by_id[oid] = v
                self._index_ids(v)

    # -- generic -----------------------------------------------------------------
    def resolve(self, value, default_id: str | None = None) -> Object:
        """Instantiate a nested plugin dict or follow a ``ref``."""
        if not isinstance(value, dict) or "type" not in value:
            raise RuntimeError(f"expected a plugin dictionary, got {value!r}")
        if value["type"] == "ref":
            rid = value["id"]
            if rid in self.by_id:
                return self.by_id[rid]
            if rid not in self.dict_by_id:
                raise RuntimeError(f"reference to unknown object id '{rid}'")
            return self.make(self.dict_by_id[rid], rid)
        return self.make(value, value.get("id", default_id))

    def make(self, d: dict, oid: str | None) -> Object:
        if oid and oid in self.by_id:
            return self.by_id[oid]
        ty = d["type"]
        if ty in _KNOWN_UNSUPPORTED:
            raise RuntimeError(f"unsupported plugin '{ty}': {_KNOWN_UNSUPPORTED[ty]}")
        kind = _KIND_OF.get(ty)
        if kind is None:
            raise RuntimeError(f"unsupported plugin type '{ty}' (no B200 kernel implementation)")
        obj = getattr(self, f"make_{kind}")(d, oid)
        if oid:
            self.by_id[oid] = obj
        return obj

    # -- leaves ------------------------------------------------------------------
    def make_texture(self, value, what: str) -> Texture:
        tex = Texture("uniform")
        tex.values["value"] = _scalar_spectrum(value, what)
        return tex

    def make_volume(self, d, oid) -> Volume:
        if not isinstance(d, dict):  # bare float -> constvolume (Properties::get_volume)
            vol = Volume("constvolume", oid)
            vol.children["value"] = self.make_texture(float(d), "constvolume.value")
            return vol
        ty = d["type"]
        vol = Volume(ty, oid)
        if ty == "constvolume":
            vol.children["value"] = self.make_texture(d.get("value", 1.0), "constvolume.value")
        elif ty == "gridvolume":
            if d.get("filter_type", "trilinear") != "nearest":
                raise RuntimeError("gridvolume: only filter_type='nearest' is supported")
            if "grid" in d:
                data = to_grid(d["grid"])
            elif "data" in d:
                data = to_grid(d["data"])
            else:
                raise RuntimeError("gridvolume: needs a 'grid' (file loading is not supported)")
            if data.shape[-1] != 1:
                raise RuntimeError("gridvolume: only 1-channel grids are supported")
            if sum(1 for s in data.shape[:3] if s > 1) > 1:
                raise RuntimeError(
                    f"gridvolume: only 1D profiles are supported, got shape {data.shape}"
                )
            vol.values["data"] = data.copy()
            vol.to_world = to_matrix(d.get("to_world"))
        elif ty == "sphericalcoordsvolume":
            vol.rmin = float(d.get("rmin", 0.0))
            vol.rmax = float(d.get("rmax", 1.0))
            if d.get("fillmin", 0.0) != 0.0 or d.get("fillmax", 0.0) != 0.0:
                raise RuntimeError("sphericalcoordsvolume: non-zero fill values are unsupported")
            vol.to_world = to_matrix(d.get("to_world"))
            vol.children["volume"] = self.make_volume(d.get("volume", 1.0), None)
            inner = vol.children["volume"]
            if inner.type == "gridvolume":
                shp = inner.values["data"].shape
                if shp[0] != 1 or shp[1] != 1:
                    raise RuntimeError(
                        "sphericalcoordsvolume: nested grid must vary along r only "
                        f"(shape [1,1,N,1]), got {shp}"
                    )
        return vol

    def make_phase(self, d, oid) -> PhaseFunction:
        ty = d["type"]
        ph = PhaseFunction(ty, oid)
        if ty == "hg":
            ph.values["g"] = float(d.get("g", 0.8))
        elif ty == "tabphase_polarized":
            # ERP/phase/tabphase_polarized.cpp:228-296: m11 on irregular nodes + m12, m22, m33, m34, m44
            ph.values["nodes"] = _parse_floats(d["nodes"], "nodes").astype(np.float32)
            ph.values["m11"] = _parse_floats(d["m11"], "m11").astype(np.float32)
            n = ph.values["nodes"].size
            if ph.values["m11"].size != n:
                raise RuntimeError(
                    "TabulatedPolarizedPhaseFunction: 'cos_theta_str' and 'm11_str' parameters "
                    "must have the same size!")
            for name in ("m12", "m22", "m33", "m34", "m44"):
                raw = d.get(name, "")
                arr = _parse_floats(raw, name).astype(np.float32) if raw != "" else np.zeros(n, np.float32)
                if arr.size != n:
                    raise RuntimeError(
                        "TabulatedPolarizedPhaseFunction: the provided parameters must have the "
                        "same size as 'cos_theta_str'!")
                ph.values[name] = arr
        elif ty in ("rayleigh", "rayleigh_polarized"):
            dep = d.get("depolarization", 0.0)  # rayleigh.cpp:48 / rayleigh_polarized.cpp: a volume (get_volume)
            # a per-layer profile is carried as a blend of two constant-depolarization leaves (phase_leaves)
            ph.children["depolarization"] = self.make_volume(dep, None)
        elif ty == "tabphase":
            ph.values["values"] = _parse_floats(d["values"], "tabphase.values").astype(np.float32)
        elif ty == "tabphase_irregular":
            ph.values["values"] = _parse_floats(d["values"], "values").astype(np.float32)
            ph.values["nodes"] = _parse_floats(d["nodes"], "nodes").astype(np.float32)
            n, v = ph.values["nodes"], ph.values["values"]
            if n.size != v.size:
                raise RuntimeError("'nodes' and 'values' must have the same length")
            if n[0] != -1.0 or n[-1] != 1.0:
                raise RuntimeError(f"'nodes' bounds must be [-1, 1], got [{n[0]}, {n[-1]}]")
        elif ty == "blendphase":
            nested = [v for k, v in d.items() if isinstance(v, dict) and k not in ("weight",)
                      and _KIND_OF.get(v.get("type")) == "phase" or
                      (isinstance(v, dict) and v.get("type") == "ref" and k != "weight")]
            if len(nested) != 2:
                raise RuntimeError("BlendPhase: Two child phase functions must be specified!")
            ph.children["weight"] = self.make_volume(d.get("weight", 0.5), None)
            ph.children["phase_0"] = self.resolve(nested[0])
            ph.children["phase_1"] = self.resolve(nested[1])
        elif ty == "multiphase":
            # ERP/phase/multiphase.cpp:75-112: nested phase functions in order of appearance, `weight<i>` volumes
            # (normalised internally), optional mixture MIS of the sampling weight (default on)
            ph.use_mis = bool(d.get("use_mis", True))
            nested = [v for k, v in d.items() if isinstance(v, dict) and not k.startswith("weight")
                      and (_KIND_OF.get(v.get("type")) == "phase" or v.get("type") == "ref")]
            if len(nested) < 2:
                raise RuntimeError("MultiPhase: At least 2 child phase functions must be specified!")
            for i, child in enumerate(nested):
                if f"weight{i}" not in d:
                    raise RuntimeError(f"multiphase: missing required parameter 'weight{i}'")
                ph.children[f"phase{i}"] = self.resolve(child)
                ph.children[f"weight{i}"] = self.make_volume(d[f"weight{i}"], None)
        return ph

    def make_bsdf(self, d, oid) -> BSDF:
        ty = d["type"]
        b = BSDF(ty, oid)
        tex = lambda name, default: self.make_texture(d.get(name, default), f"{ty}.{name}")  # noqa: E731
        if ty == "diffuse":
            b.children["reflectance"] = tex("reflectance", 0.5)
        elif ty == "rpv":  # rpv.cpp:80-92
            b.children["rho_0"] = tex("rho_0", 0.1)
            b.children["g"] = tex("g", 0.0)
            b.children["k"] = tex("k", 0.1)
            b.rho_c_tied = "rho_c" not in d  # rpv.cpp:84-87: rho_c follows rho_0 unless given
            if not b.rho_c_tied:
                b.children["rho_c"] = tex("rho_c", d["rho_c"])
        elif ty == "rtls":  # rtls.cpp:63-78
            b.children["f_iso"] = tex("f_iso", 0.209741)
            b.children["f_vol"] = tex("f_vol", 0.081384)
            b.children["f_geo"] = tex("f_geo", 0.004140)
            b.h, b.r, b.b = float(d.get("h", 2.0)), float(d.get("r", 1.0)), float(d.get("b", 1.0))
        elif ty == "hapke":
            for name in ("w", "b", "c", "theta", "B_0", "h"):
                if name not in d:
                    raise RuntimeError(f"hapke: missing required parameter '{name}'")
                b.children[name] = tex(name, 0.0)
        elif ty == "selectbsdf":
            # ERP/bsdfs/selectbsdf.cpp:62-96: an index texture picks one of >= 2 nested BSDFs; traverse() publishes
            # `indices` and `bsdf_<i>` in order of appearance.  A 1D scene has no surface position, so the index
            # texture must be uniform: the selected BSDF is then THE surface BSDF (resolved by the flattener; an
            # update of `indices.value` re-creates the device scene).
            nested = [v for k, v in d.items() if isinstance(v, dict) and k != "indices"
                      and (_KIND_OF.get(v.get("type")) == "bsdf" or v.get("type") == "ref")]
            if len(nested) < 2:
                raise RuntimeError("SelectBSDF: At least two child BSDFs must be specified!")
            idx = d.get("indices")
            if isinstance(idx, dict) and idx.get("type") != "uniform":
                raise RuntimeError(f"selectbsdf: only a uniform 'indices' texture is supported, got '{idx.get('type')}' "
                                   "(the analytic 1D surfaces carry no texture coordinates)")
            if idx is None:
                raise RuntimeError("selectbsdf: missing required parameter 'indices'")
            b.children["indices"] = tex("indices", 0.0)
            b.children["indices"].rebuild_on_change = True
            for i, child in enumerate(nested):
                b.children[f"bsdf_{i}"] = self.resolve(child)
        elif ty == "blendbsdf":
            # MI/src/bsdfs/blendbsdf.cpp as emitted by CentralPatchSurface (scenes/surface/_central_patch.py:
            # 185-215): weight = the 3x3 central-patch mask (nearest filter, clamped), so the blend is a switch:
            # bsdf_1 on the patch, bsdf_0 around it. Any other weight texture is rejected.
            wt = d.get("weight")
            ok = (isinstance(wt, dict) and wt.get("type") == "bitmap"
                  and str(wt.get("filename", "")).endswith("central_patch_surface_mask.bmp")
                  and wt.get("filter_type", "bilinear") == "nearest" and wt.get("wrap_mode", "repeat") == "clamp")
            if not ok:
                raise RuntimeError("blendbsdf: only the CentralPatchSurface form is supported (weight = the "
                                   "'central_patch_surface_mask.bmp' bitmap, nearest filter, clamped)")
            m = to_matrix(wt.get("to_uv"))
            if not (m[0, 0] > 0 and m[1, 1] > 0 and abs(m[0, 1]) < 1e-12 and abs(m[1, 0]) < 1e-12):
                raise RuntimeError("blendbsdf: unsupported 'to_uv' transform of the patch mask")
            b.uv_scale = (float(m[0, 0]), float(m[1, 1]))
            for key in ("bsdf_0", "bsdf_1"):
                if key not in d:
                    raise RuntimeError("BlendBSDF: Two child BSDFs must be specified!")
                child = self.resolve(d[key])
                if child.type not in ("diffuse", "rpv", "rtls", "hapke"):
                    raise RuntimeError(f"blendbsdf: nested BSDF '{child.type}' is not supported")
                b.children[key] = child
        elif ty == "bilambertian":  # ERP/bsdfs/bilambertian.cpp:44-49
            b.children["reflectance"] = tex("reflectance", 0.5)
            b.children["transmittance"] = tex("transmittance", 0.5)
        elif ty == "ocean_legacy":
            b.values["wavelength"] = float(d.get("wavelength", 550.0))
            b.values["wind_speed"] = float(d.get("wind_speed", 0.1))
            b.values["wind_direction"] = float(d.get("wind_direction", 0.0))
            b.values["chlorinity"] = float(d.get("chlorinity", 19.0))
            b.values["pigmentation"] = float(d.get("pigmentation", 0.3))
            b.values["shadowing"] = bool(d.get("shadowing", True))
            # ocean_legacy.cpp:299, :360-361: the derived whitecap coverage is published too (read-only: the
            # kernel re-derives it from the wind speed at every parameter update)
            b.values["coverage"] = float(min(max(2.95e-06 * b.values["wind_speed"] ** 3.52, 0.0), 1.0))
            b.component = int(d.get("component", 0))
            if b.component != 0:
                raise RuntimeError("ocean_legacy: only component=0 (full BRDF) is supported")
        elif ty == "ocean_mishchenko":  # ocean_mishchenko.cpp:95-113 (`shadowing` is read and never used)
            b.values["wind_speed"] = float(d.get("wind_speed", 0.1))
            b.children["eta"] = tex("eta", 1.33)
            b.children["k"] = tex("k", 0.0)
            b.children["ext_ior"] = tex("ext_ior", 1.000277)
        elif ty == "ocean_grasp":  # ocean_grasp.cpp:120-146
            if "wavelength" not in d:
                raise RuntimeError("ocean_grasp: missing required parameter 'wavelength'")
            b.values["wavelength"] = float(d["wavelength"])
            b.children["wind_speed"] = tex("wind_speed", 0.1)
            b.children["eta"] = tex("eta", 1.33)
            b.children["k"] = tex("k", 0.0)
            b.children["ext_ior"] = tex("ext_ior", 1.000277)
            b.children["water_body_reflectance"] = tex("water_body_reflectance", 0.0)
            b.component = int(d.get("component", 0))
            if b.component != 0:
                raise RuntimeError("ocean_grasp: only component=0 (full BRDF) is supported")
        elif ty == "mqdiffuse":  # mqdiffuse.cpp:60-92
            if "grid" in d and "filename" in d:
                raise RuntimeError('Cannot specify both "grid" and "filename".')
            if "grid" not in d:
                raise RuntimeError("mqdiffuse: needs a 'grid' (file loading is not supported)")
            data = to_grid(d["grid"])
            if data.shape[-1] != 1:
                raise RuntimeError("mqdiffuse: only 1-channel grids are supported")
            b.table = np.ascontiguousarray(data[..., 0], dtype=np.float32)  # [z, y, x] = [cos_theta_i, phi_d, cos_theta_o]
        elif ty == "measured_mono":  # measured_mono.cpp:47-54, :224-226
            from ._measured import MeasuredData

            if "filename" not in d:
                raise RuntimeError('Property "filename" has not been specified!')
            b.measured = MeasuredData(str(d["filename"]))
            b.values["wavelength"] = float(d.get("wavelength", 550.0))
            b.rebuild_on_change = True  # the table is re-blended at the new wavelength and uploaded again
        elif ty == "maignan":  # maignan.cpp:92-101 (constructor defaults, not the documented ones)
            b.children["C"] = tex("C", 0.1)
            b.children["ndvi"] = tex("ndvi", 0.0)
            b.children["refr_re"] = tex("refr_re", 1.5)
            b.children["refr_im"] = tex("refr_im", 0.0)
            b.children["ext_ior"] = tex("ext_ior", 1.000277)
        return b

    def make_medium(self, d, oid) -> Medium:
        ty = d["type"]
        m = Medium(ty, oid)
        if "phase" in d:
            m.children["phase_function"] = self.resolve(d["phase"])
        else:
            cands = [v for v in d.values() if isinstance(v, dict)
                     and (_KIND_OF.get(v.get("type")) == "phase")]
            m.children["phase_function"] = (
                self.resolve(cands[0]) if cands else PhaseFunction("isotropic")
            )
        m.children["albedo"] = self.make_volume(d.get("albedo", 0.75), None)
        m.children["sigma_t"] = self.make_volume(d.get("sigma_t", 1.0), None)
        m.values["scale"] = float(d.get("scale", 1.0))
        if not d.get("has_spectral_extinction", True):
            raise RuntimeError("has_spectral_extinction=False is not supported")
        if not d.get("sample_emitters", True):
            raise RuntimeError("sample_emitters=False is not supported")
        return m

    def make_emitter(self, d, oid) -> Emitter:
        e = Emitter(d["type"], oid)
        if "direction" in d:
            if "to_world" in d:
                raise RuntimeError("Only one of the parameters 'direction' and 'to_world' can be specified at the "
                                   "same time!'")
            direction = np.asarray(d["direction"], dtype=np.float64)
        else:
            direction = to_matrix(d.get("to_world"))[:3, :3] @ np.array([0.0, 0.0, 1.0])
        e.direction = direction / np.linalg.norm(direction)
        e.angular_diameter = 0.0
        if d["type"] == "astroobject":
            # astroobject.cpp:62-88: `direction` / to_world * z points TOWARDS the object (the opposite of
            # `directional`, whose direction is the one light travels); default diameter = the Sun's
            e.direction = -e.direction
            e.angular_diameter = float(d.get("angular_diameter", 0.5358))
            if not (0.0 < e.angular_diameter < 180.0):
                raise RuntimeError("Invalid angular diameter specified! (must be in ]0, 180[°)")
        e.children["irradiance"] = self.make_texture(d.get("irradiance", 1.0), "irradiance")
        return e

    @staticmethod
    def _disk_row(m: np.ndarray) -> list[float]:
        """(centre, unit normal, radius) of a `disk` with to_world `m` (MI/src/shapes/disk.cpp:96-122)."""
        a = m[:3, :3]
        du, dv = np.linalg.norm(a[:, 0]), np.linalg.norm(a[:, 1])
        if abs(du - dv) > 1e-6 * max(du, dv) or abs(a[:, 0] @ a[:, 1]) > 1e-6 * du * dv:
            raise RuntimeError("disk: only uniformly scaled (circular) disks are supported")
        n = np.cross(a[:, 0], a[:, 1])
        n /= np.linalg.norm(n)
        return [m[0, 3], m[1, 3], m[2, 3], n[0], n[1], n[2], 0.5 * (du + dv)]

    def _child_bsdf(self, d):
        bsdf = d.get("bsdf")
        if bsdf is None:
            cands = [v for v in d.values() if isinstance(v, dict)
                     and (_KIND_OF.get(v.get("type")) == "bsdf" or
                          (v.get("type") == "ref" and v.get("id") in self.dict_by_id and
                           _KIND_OF.get(self.dict_by_id[v["id"]].get("type")) == "bsdf"))]
            bsdf = cands[0] if cands else {"type": "diffuse"}
        return self.resolve(bsdf)

    def make_shape(self, d, oid) -> Shape:
        ty = d["type"]
        s = Shape(ty, oid)
        if ty == "shapegroup":
            # src/eradiate/scenes/biosphere/_core.py:266-275: a group of `disk` leaves. The disks are
            # kept as one [n, 7] array, not as one node each (a RAMI canopy has > 10^5 of them).
            rows, trunk_rows, cyl_rows, bsdf, trunk_bsdf = [], [], [], None, None
            tri_rows, tri_bsdf, mesh_bsdfs = [], [], []
            for key, v in d.items():
                if not isinstance(v, dict) or "type" not in v:
                    continue
                if v["type"] in _KNOWN_UNSUPPORTED:
                    raise RuntimeError(f"unsupported plugin '{v['type']}': {_KNOWN_UNSUPPORTED[v['type']]}")
                if v["type"] not in ("disk", "cylinder", "ply", "obj"):
                    raise RuntimeError(f"shapegroup: unsupported child shape '{v['type']}' "
                                       "(only 'disk', 'cylinder', 'ply' and 'obj')")
                b = self._child_bsdf(v)
                if v["type"] in ("ply", "obj"):
                    # MeshTreeElement (_tree.py:440-478): a triangle mesh with its own bilambertian BSDF and a
                    # scaling to_world (mesh units -> kernel length unit)
                    from ._mesh import load_triangles

                    if b.type != "bilambertian":
                        raise RuntimeError(f"mesh canopy elements must carry a bilambertian BSDF, got '{b.type}'")
                    if "filename" not in v:
                        raise RuntimeError('Property "filename" has not been specified!')
                    tri, _ = load_triangles(v["type"], str(v["filename"]), to_matrix(v.get("to_world")),
                                            face_normals=bool(v.get("face_normals", False)),
                                            flip_normals=bool(v.get("flip_normals", False)))
                    if not any(b is m for m in mesh_bsdfs):
                        mesh_bsdfs.append(b)
                    tri_rows.append(tri)
                    tri_bsdf.append(np.full(tri.shape[0], next(i for i, m in enumerate(mesh_bsdfs) if m is b), dtype=np.int32))
                    continue
                if b.type == "bilambertian" and v["type"] == "disk":  # a leaf
                    if bsdf is not None and b is not bsdf:
                        raise RuntimeError("shapegroup: all leaves of a group must share one BSDF")
                    bsdf = b
                    rows.append(self._disk_row(to_matrix(v.get("to_world"))))
                elif b.type == "diffuse":  # trunk parts (_tree.py:150-180): a cylinder and its cap
                    if trunk_bsdf is not None and b is not trunk_bsdf:
                        raise RuntimeError("shapegroup: all trunk parts of a group must share one BSDF")
                    trunk_bsdf = b
                    if v["type"] == "disk":
                        trunk_rows.append(self._disk_row(to_matrix(v.get("to_world"))))
                    else:  # MI/src/shapes/cylinder.cpp:110-135: p0, p1, radius
                        if "to_world" in v and not np.allclose(to_matrix(v["to_world"]), np.eye(4)):
                            raise RuntimeError("cylinder: specify p0 / p1 / radius (to_world is unsupported)")
                        p0 = np.asarray(v.get("p0", [0.0, 0.0, 0.0]), dtype=np.float64)
                        p1 = np.asarray(v.get("p1", [0.0, 0.0, 1.0]), dtype=np.float64)
                        if not np.linalg.norm(p1 - p0) > 0:
                            raise RuntimeError("cylinder: p0 and p1 must differ")
                        cyl_rows.append([*p0, *p1, float(v.get("radius", 1.0))])
                else:
                    raise RuntimeError(f"canopy leaves must carry a bilambertian BSDF and trunks a diffuse one, "
                                       f"got '{b.type}' on a {v['type']}")
            if not rows and not tri_rows:
                raise RuntimeError("shapegroup: no leaf (bilambertian disk or mesh) among the child shapes")
            s.triangles = np.concatenate(tri_rows) if tri_rows else np.zeros((0, 18), dtype=np.float32)
            s.triangle_bsdf = np.concatenate(tri_bsdf) if tri_bsdf else np.zeros(0, dtype=np.int32)
            s.mesh_bsdfs = mesh_bsdfs
            for i, m in enumerate(mesh_bsdfs):
                s.children[f"mesh_bsdf_{i}"] = m
            s.disks = np.asarray(rows, dtype=np.float64).reshape(-1, 7)
            s.trunk_disks = np.asarray(trunk_rows, dtype=np.float64).reshape(-1, 7)
            s.cylinders = np.asarray(cyl_rows, dtype=np.float64).reshape(-1, 7)
            if bsdf is not None:
                s.children["bsdf"] = bsdf
            if trunk_bsdf is not None:
                s.children["trunk_bsdf"] = trunk_bsdf
            return s
        if ty in ("ply", "obj"):
            raise RuntimeError(f"'{ty}': triangle meshes are supported as canopy elements inside a shapegroup only")
        if ty == "instance":
            # _core.py:277-296: `group` reference + translation
            g = d.get("group") or next((v for v in d.values() if isinstance(v, dict) and v.get("type") in ("ref", "shapegroup")), None)
            if g is None:
                raise RuntimeError("instance: missing 'group' reference")
            grp = self.resolve(g)
            if grp.type != "shapegroup":
                raise RuntimeError("instance: 'group' must reference a shapegroup")
            m = to_matrix(d.get("to_world"))
            if not np.allclose(m[:3, :3], np.eye(3), atol=1e-12):
                raise RuntimeError("instance: only translations are supported as to_world")
            s.group = grp
            s.offset = m[:3, 3].astype(np.float64).copy()
            return s
        s.to_world = to_matrix(d.get("to_world"))
        if ty == "disk":  # a free-standing leaf: a group of one, instanced once at the origin
            bsdf = self._child_bsdf(d)
            if bsdf.type != "bilambertian":
                raise RuntimeError(f"canopy leaves must carry a bilambertian BSDF, got '{bsdf.type}'")
            s.disks = np.asarray([self._disk_row(s.to_world)], dtype=np.float64)
            s.trunk_disks = np.zeros((0, 7))
            s.cylinders = np.zeros((0, 7))
            s.triangles, s.triangle_bsdf, s.mesh_bsdfs = np.zeros((0, 18), dtype=np.float32), np.zeros(0, dtype=np.int32), []
            s.children["bsdf"] = bsdf
            return s
        if ty == "sphere":
            s.center = np.asarray(d.get("center", [0.0, 0.0, 0.0]), dtype=np.float64)
            s.radius = float(d.get("radius", 1.0))
            m = s.to_world
            scale = np.linalg.norm(m[:3, :3], axis=0)
            if not np.allclose(scale, scale[0], rtol=1e-9):
                raise RuntimeError("sphere: non-uniform scaling is unsupported")
            s.center = m[:3, :3] @ s.center + m[:3, 3]
            s.radius *= scale[0]
        bsdf = d.get("bsdf")
        if bsdf is None:
            cands = [v for v in d.values() if isinstance(v, dict)
                     and (_KIND_OF.get(v.get("type")) == "bsdf" or
                          (v.get("type") == "ref" and v.get("id") in self.dict_by_id and
                           _KIND_OF.get(self.dict_by_id[v["id"]].get("type")) == "bsdf"))]
            bsdf = cands[0] if cands else {"type": "diffuse"}
        s.children["bsdf"] = self.resolve(bsdf)
        if "interior" in d:
            s.children["interior_medium"] = self.resolve(d["interior"])
        if "exterior" in d:
            raise RuntimeError("shapes with an exterior medium are unsupported")
        return s

    def make_sensor(self, d, oid) -> Sensor:
        ty = d["type"]
        s = Sensor(ty, oid)
        film_d = d.get("film", {"type": "hdrfilm"})
        if film_d.get("type", "hdrfilm") != "hdrfilm":
            raise RuntimeError(f"unsupported film type '{film_d.get('type')}'")
        rf = film_d.get("rfilter", {"type": "box"})
        if isinstance(rf, dict) and rf.get("type", "box") != "box":
            raise RuntimeError("only the box reconstruction filter is supported")
        film = Film("hdrfilm")
        film.width = int(film_d.get("width", 768))
        film.height = int(film_d.get("height", 576))
        film.pixel_format = film_d.get("pixel_format", "rgb")
        s.children["film"] = film
        samp_d = d.get("sampler", {"type": "independent"})
        if samp_d.get("type", "independent") != "independent":
            raise RuntimeError(f"unsupported sampler '{samp_d.get('type')}' (only 'independent')")
        sampler = Sampler("independent")
        sampler.sample_count = int(samp_d.get("sample_count", 4))
        s.children["sampler"] = sampler
        s.in_medium = False
        if "medium" in d:
            if ty not in ("perspective", "mradiancemeter"):
                raise RuntimeError("distant sensors inside a medium are unsupported")
            s.in_medium = True
        s.ray_offset = float(d.get("ray_offset", -1.0))
        s.to_world = to_matrix(d.get("to_world"))
        if ty == "perspective":
            # MI/src/sensors/perspective.cpp:130-175 + sensor.cpp parse_fov
            if "focal_length" in d:
                raise RuntimeError("perspective: specify 'fov' (focal_length is unsupported)")
            fov = float(d.get("fov", 0.0))
            if not (0.0 < fov < 180.0):
                raise RuntimeError("The horizontal field of view must be in the range [0, 180]!")
            axis = d.get("fov_axis", "x")
            aspect = film.width / film.height
            if axis == "smaller":
                axis = "y" if aspect > 1 else "x"
            elif axis == "larger":
                axis = "x" if aspect > 1 else "y"
            if axis == "y":
                fov = np.degrees(2.0 * np.arctan(np.tan(0.5 * np.radians(fov)) * aspect))
            elif axis == "diagonal":
                diag = 2.0 * np.tan(0.5 * np.radians(fov))
                fov = np.degrees(2.0 * np.arctan(0.5 * diag * aspect / np.sqrt(1.0 + aspect * aspect)))
            elif axis != "x":
                raise RuntimeError(f"perspective: invalid fov_axis '{axis}'")
            # perspective.cpp:186-194 publishes these (an update re-creates the device scene)
            s.values["x_fov"] = float(fov)
            s.values["near_clip"] = float(d.get("near_clip", 1e-2))
            s.values["far_clip"] = float(d.get("far_clip", 1e4))
            s.values["principal_point_offset_x"] = float(d.get("principal_point_offset_x", 0.0))
            s.values["principal_point_offset_y"] = float(d.get("principal_point_offset_y", 0.0))
            if s.values["principal_point_offset_x"] != 0.0 or s.values["principal_point_offset_y"] != 0.0:
                raise RuntimeError("perspective: a principal point offset is not supported")
            a = s.to_world[:3, :3]
            if not np.allclose(a.T @ a, np.eye(3), atol=1e-6):
                raise RuntimeError("Scale factors in the camera-to-world transformation are not allowed!")
        if ty == "mradiancemeter":  # ERP/sensors/mradiancemeter.cpp:72-133
            if "to_world" in d:
                raise RuntimeError(
                    "This sensor is specified through a set of origin and direction values and cannot "
                    "use the to_world transform."
                )
            org = _parse_floats(d["origins"], "origins")
            dirs = _parse_floats(d["directions"], "directions")
            if org.size % 3 != 0:
                raise RuntimeError(f"Invalid specification! Number of parameters {org.size}, is not a multiple of three.")
            if org.size != dirs.size:
                raise RuntimeError(
                    f"Invalid specification! Number of parameters for origins and directions ({org.size}, "
                    f"{dirs.size}) are not equal.")
            s.origins, s.directions = org.reshape(-1, 3), dirs.reshape(-1, 3)
            if (film.width, film.height) != (s.origins.shape[0], 1):
                raise RuntimeError(
                    f"Film size must be [n_radiancemeters, 1]. Expected [{s.origins.shape[0]}, 1], "
                    f"found: [{film.width}, {film.height}]")
        if ty == "mpdistant":  # ERP/sensors/mpdistant.cpp:171-205
            if "direction" in d:
                if "to_world" in d:
                    raise RuntimeError("Only one of the parameters 'direction' and 'to_world' can be specified at the same time!")
                v = np.asarray(d["direction"], dtype=np.float64)
                v = v / np.linalg.norm(v)
                sign = np.copysign(1.0, v[2])  # coordinate_system(direction) -> (s, t): up = t
                a, b = -1.0 / (sign + v[2]), v[0] * v[1] * (-1.0 / (sign + v[2]))
                up = np.array([b, sign + v[1] * v[1] * a, -v[1]])
                s.to_world = ScalarTransform4f().look_at([0.0, 0.0, 0.0], v, up).matrix
            if float(d.get("target_radius", -1.0)) >= 0.0:
                raise RuntimeError("mpdistant: 'target_radius' is unsupported")
        if ty == "mdistant":
            if "to_world" in d:
                raise RuntimeError(
                    "This sensor is specified through a set of origin and direction "
                    "values and cannot use the to_world transform."
                )
            dirs = _parse_floats(d["directions"], "directions")
            if dirs.size % 3 != 0:
                raise RuntimeError(
                    f"Invalid specification! Number of parameters {dirs.size}, is not a "
                    "multiple of three."
                )
            s.directions = dirs.reshape(-1, 3)
            if (film.width, film.height) != (s.directions.shape[0], 1):
                raise RuntimeError(
                    f"Film size must be [sensor_count, 1]. Expected "
                    f"[{s.directions.shape[0]}, 1], got [{film.width}, {film.height}]"
                )
        tgt = d.get("target")
        s.target_to_world = np.eye(4)
        s.target_point = np.zeros(3)
        if tgt is None:
            s.target_type = _abi.TARGET_NONE
        elif isinstance(tgt, dict):
            if tgt.get("type") == "rectangle":
                s.target_type = _abi.TARGET_RECTANGLE
            elif tgt.get("type") == "disk":
                s.target_type = _abi.TARGET_DISK
            else:
                raise RuntimeError(f"unsupported target shape '{tgt.get('type')}'")
            s.target_to_world = to_matrix(tgt.get("to_world"))
        else:
            s.target_type = _abi.TARGET_POINT
            s.target_point = np.asarray(tgt, dtype=np.float64).reshape(3)
        return s

    def make_integrator(self, d, oid) -> Integrator:
        ty = d["type"]
        it = Integrator(ty, oid)
        it.moment = False
        it.stokes = False
        it.meridian_align = False
        inner = d
        if ty == "stokes":  # MI/src/integrators/stokes.cpp: must be the outermost wrapper
            it.stokes = True
            it.meridian_align = bool(d.get("meridian_align", False))
            nested = [v for v in d.values() if isinstance(v, dict) and "type" in v]
            if len(nested) != 1:
                raise RuntimeError("Must specify a sub-integrator!")
            inner = d = nested[0]
            ty = inner["type"]
        if ty == "moment":
            it.moment = True
            nested = [v for v in d.values() if isinstance(v, dict) and "type" in v]
            if len(nested) != 1:
                raise RuntimeError("moment: exactly one nested integrator is supported")
            inner = nested[0]
            if inner["type"] in _KNOWN_UNSUPPORTED:
                raise RuntimeError(
                    f"unsupported plugin '{inner['type']}': {_KNOWN_UNSUPPORTED[inner['type']]}"
                )
            if inner["type"] not in ("volpath", "volpathmis", "piecewise_volpath", "path"):
                raise RuntimeError(f"unsupported nested integrator '{inner['type']}'")
        it.kernel_type = inner["type"]
        it.max_depth = int(inner.get("max_depth", -1))
        it.rr_depth = int(inner.get("rr_depth", 5))
        it.hide_emitters = bool(inner.get("hide_emitters", False))  # MI/src/render/integrator.cpp:29
        if it.max_depth < 0 and it.max_depth != -1:
            raise RuntimeError(
                '"max_depth" must be set to -1 (infinite) or a value >= 0'
            )
        if it.rr_depth <= 0:
            raise RuntimeError('"rr_depth" must be set to a value greater than zero!')
        return it

    # -- scene ---------------------------------------------------------------------
    def load(self) -> Object:
        root = self.root
        if root.get("type") != "scene":
            return self.make(root, root.get("id"))
        scene = Scene()
        scene._integrator_key = None
        for key, value in root.items():
            if not isinstance(value, dict) or "type" not in value:
                continue
            if value["type"] == "ref":
                continue
            obj = self.make(value, value.get("id", key))
            if obj._id == "":
                obj._id = key
            scene.children[obj.id() or key] = obj  # every top-level object is a child (scene.cpp:510-522)
            if obj.plugin_kind == "integrator":
                scene._integrator_key = obj.id() or key
            elif obj.plugin_kind == "sensor":
                scene._sensors.append(obj)
        if scene._integrator_key is None:
            raise RuntimeError("scene has no integrator")
        scene.flat  # validate eagerly: load errors surface at mi_load_dict time
        return scene


def load_dict(d: dict) -> Object:
    return _Loader(d).load()


# ------------------------------------------------------------------------------
#                                  flattener
# ------------------------------------------------------------------------------


class FlatScene:
    """Flat view of a :class:`Scene`; owns the numpy buffers the C ABI points to."""

    def __init__(self, scene: Scene):
        self.scene = scene
        self._keepalive: list = []
        self._extract()

    # -- extraction ----------------------------------------------------------------
    def _extract(self) -> None:
        sc = self.scene
        shapes = [o for o in sc.children.values() if isinstance(o, Shape)]
        emitters = [o for o in sc.children.values() if isinstance(o, Emitter)]
        if len(emitters) != 1:
            raise RuntimeError(f"exactly one directional emitter is supported, got {len(emitters)}")
        self.emitter = emitters[0]
        self.integrator = sc.integrator()

        canopy = [s for s in shapes if s.type in ("shapegroup", "instance", "disk")]
        shapes = [s for s in shapes if s.type not in ("shapegroup", "instance", "disk")]
        atm = [s for s in shapes if "interior_medium" in s.children]
        srf = [s for s in shapes if "interior_medium" not in s.children]
        if len(atm) > 1:
            raise RuntimeError("at most one medium-bearing shape is supported (1D atmospheres)")
        if len(srf) != 1:
            raise RuntimeError(f"exactly one surface shape is supported, got {len(srf)}")
        self.surface_shape = srf[0]
        self.atm_shape = atm[0] if atm else None
        self.bsdf: BSDF = self.surface_shape.children["bsdf"]
        if self.bsdf.type == "selectbsdf":  # selectbsdf.cpp:98-110: UInt32(indices->eval_1(si)) picks the BSDF
            n_sel = sum(1 for k in self.bsdf.children if k.startswith("bsdf_"))
            i_sel = int(np.uint32(self.bsdf.children["indices"].values["value"]))
            if not 0 <= i_sel < n_sel:
                raise RuntimeError(f"selectbsdf: index {i_sel} out of range (the plugin holds {n_sel} BSDFs)")
            self.bsdf = self.bsdf.children[f"bsdf_{i_sel}"]
            if self.bsdf.type in ("selectbsdf", "blendbsdf"):
                raise RuntimeError("selectbsdf: nested BSDF adapters are not supported")
        self.patch_bsdf: BSDF | None = None
        if self.bsdf.type == "blendbsdf":  # CentralPatchSurface: background + patch
            self.patch_blend = self.bsdf
            self.patch_bsdf = self.bsdf.children["bsdf_1"]
            self.bsdf = self.bsdf.children["bsdf_0"]
        if self.bsdf.type == "null":
            raise RuntimeError("the surface shape must not carry a null BSDF")
        if self.atm_shape is not None and self.atm_shape.children["bsdf"].type != "null":
            raise RuntimeError("the atmosphere stencil shape must carry a null BSDF")

        # geometry ------------------------------------------------------------------
        s = self.surface_shape
        if s.type == "sphere" and self.patch_bsdf is not None:
            raise RuntimeError("CentralPatchSurface is supported in plane-parallel scenes only")
        if s.type == "sphere":
            self.geometry = _abi.GEOM_SPHERICAL_SHELL
            if not np.allclose(s.center, 0.0, atol=1e-6 * s.radius):
                raise RuntimeError("spherical-shell scenes must be centred at the origin")
            self.surface_z = s.radius
        elif s.type in ("rectangle", "arectangle"):
            self.geometry = _abi.GEOM_PLANE_PARALLEL
            n = s.to_world[:3, :3] @ np.array([0.0, 0.0, 1.0])
            if not np.allclose(n / np.linalg.norm(n), [0, 0, 1], atol=1e-9):
                raise RuntimeError("the ground rectangle must be horizontal (normal +Z)")
            self.surface_z = float(s.to_world[2, 3])
            if self.patch_bsdf is not None:
                # uv = (local + 1) / 2 over the rectangle; the mask's central third, shrunk by the to_uv scale
                a = s.to_world[:3, :3]
                if abs(a[0, 1]) > 1e-9 * abs(a[0, 0]) or abs(a[1, 0]) > 1e-9 * abs(a[1, 1]):
                    raise RuntimeError("CentralPatchSurface: the ground rectangle must be axis-aligned")
                sx, sy = self.patch_blend.uv_scale
                self.patch_rect = (float(s.to_world[0, 3]), float(s.to_world[1, 3]),
                                   abs(float(a[0, 0])) / (3.0 * sx), abs(float(a[1, 1])) / (3.0 * sy))
        else:
            raise RuntimeError(f"unsupported surface shape '{s.type}'")

        bbox_lo, bbox_hi = _shape_bbox(s)
        self.medium = None
        self.medium_top = self.surface_z
        self.medium_bottom = self.surface_z
        if self.atm_shape is not None:
            a = self.atm_shape
            lo, hi = _shape_bbox(a)
            bbox_lo, bbox_hi = np.minimum(bbox_lo, lo), np.maximum(bbox_hi, hi)
            self.medium = a.children["interior_medium"]
            if self.geometry == _abi.GEOM_SPHERICAL_SHELL:
                if a.type != "sphere":
                    raise RuntimeError("spherical-shell atmosphere must use a sphere stencil")
                self.medium_top = a.radius
            else:
                if a.type != "cube":
                    raise RuntimeError("plane-parallel atmosphere must use a cube stencil")
                self.medium_top = float(hi[2])
            self._extract_medium()
        # canopy: leaf groups + translated instances -----------------------------
        self.leaf_groups: list[Shape] = []
        self.instances: list[tuple[int, np.ndarray]] = []
        for c in canopy:
            grp = c.group if c.type == "instance" else c
            if grp not in self.leaf_groups:
                self.leaf_groups.append(grp)
            if c.type != "shapegroup":  # a shapegroup alone is not rendered (MI/src/shapes/shapegroup.cpp)
                off = c.offset if c.type == "instance" else np.zeros(3)
                self.instances.append((self.leaf_groups.index(grp), off))
        if self.instances:
            if self.geometry != _abi.GEOM_PLANE_PARALLEL:
                raise RuntimeError("explicit canopies are supported in plane-parallel scenes only "
                                   "(src/eradiate/experiments/_canopy_atmosphere.py:74)")
            for gi, off in self.instances:
                g = self.leaf_groups[gi]
                dk = np.vstack([g.disks, g.trunk_disks])
                ext = dk[:, 6:7] * np.sqrt(np.maximum(1.0 - dk[:, 3:6] ** 2, 0.0))
                lo, hi = np.full(3, np.inf), np.full(3, -np.inf)
                if dk.shape[0]:
                    lo, hi = (dk[:, :3] - ext).min(axis=0) + off, (dk[:, :3] + ext).max(axis=0) + off
                for c in g.cylinders:
                    lo = np.minimum(lo, np.minimum(c[:3], c[3:6]) - c[6] + off)
                    hi = np.maximum(hi, np.maximum(c[:3], c[3:6]) + c[6] + off)
                if g.triangles.shape[0]:
                    v = g.triangles[:, :9].reshape(-1, 3).astype(np.float64)
                    lo, hi = np.minimum(lo, v.min(axis=0) + off), np.maximum(hi, v.max(axis=0) + off)
                    if v[:, 2].min() + off[2] < self.surface_z - 1e-6:
                        raise RuntimeError("canopy meshes must lie above the ground surface")
                dk = g.disks
                bbox_lo, bbox_hi = np.minimum(bbox_lo, lo), np.maximum(bbox_hi, hi)
                if dk.shape[0] and (dk[:, 2] + off[2]).min() < self.surface_z:
                    raise RuntimeError("canopy leaf centres must lie above the ground surface")
                if self.atm_shape is not None and hi[2] > self.medium_top:
                    raise RuntimeError("canopy leaves must lie below the top of the atmosphere")
        self.bsphere_center = 0.5 * (bbox_lo + bbox_hi)
        self.bsphere_radius = float(0.5 * np.linalg.norm(bbox_hi - bbox_lo))
        self.sensors = sc.sensors()
        if not self.sensors:
            raise RuntimeError("scene has no sensor")
        if self.integrator.kernel_type == "path" and self.medium is not None:
            raise RuntimeError("the 'path' integrator ignores participating media; use volpath")
        for sn in self.sensors:
            if sn.type == "perspective" and self.geometry != _abi.GEOM_PLANE_PARALLEL:
                raise RuntimeError("perspective sensors are supported in plane-parallel scenes only")
        # Polarized transport = the reference's *_polarized variants.  The variant is global state in
        # Mitsuba; here it is inferred from the scene: a `stokes` integrator or a polarized phase plugin.
        self.polarized = bool(getattr(self.integrator, "stokes", False))
        if self.medium is not None:
            def has_pol(ph):
                if ph.type == "blendphase":
                    return has_pol(ph.children["phase_0"]) or has_pol(ph.children["phase_1"])
                return ph.type.endswith("_polarized")
            self.polarized = self.polarized or has_pol(self.medium.children["phase_function"])
        forced = getattr(sc, "_force_polarized", None)
        if forced is not None:
            self.polarized = bool(forced)
        if self.polarized and self.integrator.kernel_type == "volpathmis":
            # volpathmis.cpp:130-132
            raise RuntimeError("This integrator currently does not support polarized mode!")


    def _extract_medium(self) -> None:
        m = self.medium
        st, al = m.children["sigma_t"], m.children["albedo"]
        self.homogeneous = m.type == "homogeneous"
        if self.integrator.kernel_type == "piecewise_volpath" and m.type != "piecewise":
            # MI/src/render/medium.cpp:99-118: only ERP/media/piecewise.cpp overrides the *_real interface
            cls = "HomogeneousMedium" if self.homogeneous else "HeterogeneousMedium"
            raise RuntimeError(f"{cls}::sample_interaction_real(): not implemented!")
        if m.type == "piecewise":
            # piecewise.cpp:445-456: a stack of horizontal layers, grid shape [1, 1, N] along z
            if self.geometry == _abi.GEOM_SPHERICAL_SHELL:
                raise RuntimeError("PiecewiseMedium: plane-parallel layer stacks only "
                                   "(src/eradiate/experiments/_helpers.py:127-165)")
            if st.type == "gridvolume" and (st.values["data"].shape[1] != 1 or st.values["data"].shape[2] != 1):
                raise RuntimeError("PiecewiseMedium: x or y resolution bigger than one, assumed shape is [1,1,n]")
        if self.homogeneous:
            if st.type != "constvolume" or al.type != "constvolume":
                raise RuntimeError("homogeneous medium expects constant sigma_t / albedo")
            self.medium_bottom = self.surface_z
            return
        top_tol = 1e-6 * abs(self.medium_top)
        if self.geometry == _abi.GEOM_SPHERICAL_SHELL:
            for v in (st, al):
                if v.type == "constvolume":
                    continue
                if v.type != "sphericalcoordsvolume":
                    raise RuntimeError(
                        "spherical-shell media must use sphericalcoordsvolume (or constant) volumes"
                    )
                scale = np.linalg.norm(v.to_world[:3, :3], axis=0)
                if not np.allclose(scale, self.medium_top, atol=top_tol) or v.rmax != 1.0:
                    raise RuntimeError("sphericalcoordsvolume extent must match the TOA sphere")
            rmin = st.rmin if st.type == "sphericalcoordsvolume" else self.surface_z / self.medium_top
            self.medium_bottom = float(rmin * self.medium_top)
        else:
            ref = next((v for v in (st, al) if v.type == "gridvolume"), None)
            if ref is None:
                self.medium_bottom = self.surface_z
            else:
                z0 = float((ref.to_world @ np.array([0.0, 0.0, 0.0, 1.0]))[2])
                z1 = float((ref.to_world @ np.array([0.0, 0.0, 1.0, 1.0]))[2])
                if abs(z1 - self.medium_top) > max(top_tol, 1e-6):
                    raise RuntimeError("gridvolume z-extent must end at the top of the slab")
                if ref.values["data"].shape[1] != 1 or ref.values["data"].shape[2] != 1:
                    raise RuntimeError("plane-parallel grids must vary along z only ([N,1,1,1])")
                self.medium_bottom = z0

    # -- per-layer arrays ----------------------------------------------------------------
    def n_layers(self) -> int:
        if self.medium is None:
            return 0
        n = 1
        for v in self._layer_volumes():
            n = max(n, v.layer_values().size)
        return n

    def _layer_volumes(self) -> list[Volume]:
        vols = [self.medium.children["sigma_t"], self.medium.children["albedo"]]

        def walk(ph):
            if ph.type == "blendphase":
                vols.append(ph.children["weight"])
                walk(ph.children["phase_0"])
                walk(ph.children["phase_1"])

        walk(self.medium.children["phase_function"])
        return vols

    def _profile(self, vol: Volume, n: int, what: str) -> np.ndarray:
        v = vol.layer_values()
        if v.size == 1:
            return np.full(n, v[0], dtype=np.float32)
        if v.size != n:
            raise RuntimeError(
                f"{what}: all per-layer volumes must share one vertical grid "
                f"({v.size} != {n} layers)"
            )
        return v.astype(np.float32)

    def phase_leaves(self, n: int):
        """Flatten the blendphase tree: list of (leaf node, probability[n])."""
        leaves: list[tuple[PhaseFunction, np.ndarray]] = []
        self._phase_mis = False

        def walk(ph: PhaseFunction, prob: np.ndarray):
            if ph.type == "blendphase":
                w = np.clip(self._profile(ph.children["weight"], n, "blendphase.weight"), 0.0, 1.0)
                walk(ph.children["phase_0"], prob * (1.0 - w))
                walk(ph.children["phase_1"], prob * w)
            elif ph.type == "multiphase":
                # multiphase.cpp:123-207: component i is drawn with probability w_i / sum(w) and, without MIS,
                # returns its own weight -- the flattened blend.  With MIS the weight becomes
                # sum_j w_j value_j / sum_j w_j pdf_j at the sampled direction, which is the component's own
                # weight (1) whenever every component's value equals its pdf; the kernels do not carry the general
                # form (Mueller-valued or depolarized-Rayleigh components).
                k = sum(1 for name in ph.children if name.startswith("phase"))
                ws = [np.asarray(self._profile(ph.children[f"weight{i}"], n, f"multiphase.weight{i}"), np.float64)
                      for i in range(k)]
                total = np.sum(ws, axis=0)
                if np.any(total <= 0.0):
                    raise RuntimeError("multiphase: the weights must have a positive sum in every layer")
                if getattr(ph, "use_mis", True):
                    for leaf in _phase_leaf_nodes(ph):
                        dep = bool(np.any(leaf.children["depolarization"].layer_values() != 0.0)) \
                            if "depolarization" in leaf.children else False
                        if leaf.type in ("rayleigh_polarized", "tabphase_polarized") or \
                                (leaf.type == "rayleigh" and dep):
                            # the mixture weight differs from the leaf's own: carried by the kernels (`phase_mis`)
                            # when the multiphase node is the medium's phase function and its components are leaves
                            flat_children = all(ph.children[f"phase{i}"].type not in ("blendphase", "multiphase")
                                                for i in range(k))
                            if ph is not self.medium.children["phase_function"] or not flat_children:
                                raise RuntimeError(
                                    "multiphase: use_mis=True over components whose value differs from their pdf "
                                    f"(got '{leaf.type}') is only supported for a non-nested multiphase node; "
                                    "pass use_mis=False")
                            self._phase_mis = True
                for i in range(k):
                    walk(ph.children[f"phase{i}"], prob * (ws[i] / total))
            elif ph.type in ("rayleigh", "rayleigh_polarized") and \
                    ph.children["depolarization"].layer_values().size > 1:
                # rayleigh.cpp:48,79 / rayleigh_polarized.cpp: `depolarization` is a volume evaluated at the
                # interaction.  Every entry of the (Mueller) value is affine in (1, rho) / (2 + rho) -- r1 r2 =
                # 2 (1 + rho) / (2 + rho), r1 = 2 (1 - rho) / (2 + rho), r1 r3 = 2 (1 - 2 rho) / (2 + rho) -- and the
                # sampling density (1 + cos^2) does not depend on rho, so the layer's phase function IS the blend
                # of the two constant-rho leaves at the profile's extremes with the per-layer weight
                # w_hi = [b(rho) - b(lo)] / [b(hi) - b(lo)], b(rho) = rho / (2 + rho): same value, same pdf.
                # (Always two leaves for a gridded profile, so that an update cannot change the leaf count.)
                rho = self._profile(ph.children["depolarization"], n, "rayleigh.depolarization").astype(np.float64)
                lo, hi = float(rho.min()), float(rho.max())
                if hi >= 1.0:
                    raise RuntimeError("Depolarization factor must be in [0, 1[")
                b = lambda r: r / (2.0 + r)  # noqa: E731
                w_hi = (b(rho) - b(lo)) / (b(hi) - b(lo)) if hi > lo else np.zeros(n)
                leaves.append((_RayleighLeaf(ph.type, lo), (prob * (1.0 - w_hi)).astype(np.float32)))
                leaves.append((_RayleighLeaf(ph.type, hi), (prob * w_hi).astype(np.float32)))
            else:
                leaves.append((ph, prob.astype(np.float32)))

        walk(self.medium.children["phase_function"], np.ones(n, dtype=np.float32))
        if len(leaves) > _abi.MAX_PHASE:
            raise RuntimeError(
                f"phase function tree has {len(leaves)} leaves; at most {_abi.MAX_PHASE} supported"
            )
        return leaves

    def bsdf_params(self, b: BSDF | None = None) -> np.ndarray:
        b = self.bsdf if b is None else b
        p = np.zeros(_abi.MAX_BSDF_PARAMS, dtype=np.float32)
        tv = lambda name: b.children[name].values["value"]  # noqa: E731
        if b.type == "diffuse":
            p[0] = tv("reflectance")
        elif b.type == "rpv":
            p[0], p[1], p[2] = tv("rho_0"), tv("k"), tv("g")
            p[3] = tv("rho_0") if getattr(b, "rho_c_tied", False) else tv("rho_c")
        elif b.type == "rtls":
            p[0], p[1], p[2] = tv("f_iso"), tv("f_vol"), tv("f_geo")
            p[3], p[4], p[5] = b.h, b.r, b.b
        elif b.type == "hapke":
            for i, name in enumerate(("w", "b", "c", "theta", "B_0", "h")):
                p[i] = tv(name)
        elif b.type == "ocean_legacy":
            v = b.values
            p[0], p[1], p[2] = v["wavelength"], v["wind_speed"], v["wind_direction"]
            p[3], p[4], p[5] = v["chlorinity"], v["pigmentation"], float(v["shadowing"])
            p[6] = float(b.component)
        elif b.type == "ocean_mishchenko":
            p[0] = b.values["wind_speed"]
            p[1], p[2], p[3] = tv("eta"), tv("k"), tv("ext_ior")
        elif b.type == "ocean_grasp":
            p[0], p[1] = b.values["wavelength"], tv("wind_speed")
            p[2], p[3], p[4], p[5] = tv("eta"), tv("k"), tv("ext_ior"), tv("water_body_reflectance")
            p[6] = float(b.component)
        elif b.type == "maignan":
            for i, name in enumerate(("C", "ndvi", "refr_re", "refr_im", "ext_ior")):
                p[i] = tv(name)
        return p

    def trunk_reflectance(self, group: int) -> float:
        b = self.leaf_groups[group].children.get("trunk_bsdf")
        return 0.0 if b is None else float(b.children["reflectance"].values["value"])

    def leaf_bsdf_params(self, group: int) -> tuple[float, float]:
        b = self.leaf_groups[group].children.get("bsdf")
        if b is None:  # a group of meshes only
            return 0.0, 0.0
        return (float(b.children["reflectance"].values["value"]),
                float(b.children["transmittance"].values["value"]))

    def mesh_bsdf_params(self, group: int) -> np.ndarray:
        """[n_mesh_bsdfs, 2]: bilambertian (reflectance, transmittance) of the group's mesh elements."""
        return np.array([[float(b.children["reflectance"].values["value"]), float(b.children["transmittance"].values["value"])]
                         for b in self.leaf_groups[group].mesh_bsdfs], dtype=np.float32).reshape(-1, 2)

    def bsdf_type(self, b: BSDF | None = None) -> int:
        return {
            "diffuse": _abi.BSDF_DIFFUSE,
            "rpv": _abi.BSDF_RPV,
            "rtls": _abi.BSDF_RTLS,
            "hapke": _abi.BSDF_HAPKE,
            "ocean_legacy": _abi.BSDF_OCEAN_LEGACY,
            "ocean_mishchenko": _abi.BSDF_OCEAN_MISHCHENKO,
            "ocean_grasp": _abi.BSDF_OCEAN_GRASP,
            "maignan": _abi.BSDF_MAIGNAN,
            "mqdiffuse": _abi.BSDF_MQDIFFUSE,
            "measured_mono": _abi.BSDF_MEASURED_MONO,
        }[(self.bsdf if b is None else b).type]

    # -- ctypes descriptor -----------------------------------------------------------------
    def build_desc(self) -> _abi.SceneDesc:
        """Build the POD descriptor. The numpy buffers it points to are owned by the returned object
        (an earlier descriptor stays valid when ``build_desc`` is called again)."""
        keep: list = []
        d = _abi.SceneDesc()
        d.abi_version = _abi.ABI_VERSION
        d.geometry = self.geometry
        d.surface_z = self.surface_z
        d.medium_bottom = self.medium_bottom
        d.medium_top = self.medium_top
        d.bsphere_center[:] = list(self.bsphere_center)
        d.bsphere_radius = self.bsphere_radius

        n = self.n_layers()
        d.has_medium = int(self.medium is not None)
        d.n_layers = n
        d.n_phase = 0
        d.sigma_t_scale = 1.0
        if self.medium is not None:
            if n > _abi.MAX_LAYERS:
                raise RuntimeError(f"too many layers ({n} > {_abi.MAX_LAYERS})")
            m = self.medium
            sig = np.ascontiguousarray(self._profile(m.children["sigma_t"], n, "sigma_t"))
            alb = np.ascontiguousarray(self._profile(m.children["albedo"], n, "albedo"))
            keep += [sig, alb]
            d.sigma_t = sig.ctypes.data_as(_abi.c_float_p)
            d.albedo = alb.ctypes.data_as(_abi.c_float_p)
            d.sigma_t_scale = m.values["scale"]
            d.homogeneous = int(self.homogeneous)
            leaves = self.phase_leaves(n)
            d.n_phase = len(leaves)
            d.phase_mis = int(getattr(self, "_phase_mis", False))
            w = np.ascontiguousarray(np.stack([p for _, p in leaves]).astype(np.float32))
            keep.append(w)
            d.phase_weight = w.ctypes.data_as(_abi.c_float_p)
            for i, (ph, _) in enumerate(leaves):
                pd = d.phase[i]
                pd.type, params, values, nodes = _phase_leaf_desc(ph)
                pd.params[:] = params
                if values is not None:
                    values = np.ascontiguousarray(values, dtype=np.float32)
                    if values.size > _abi.MAX_PHASE_NODES:
                        raise RuntimeError("tabulated phase function has too many nodes")
                    keep.append(values)
                    pd.n_nodes = values.size
                    pd.values = values.ctypes.data_as(_abi.c_float_p)
                if nodes is not None:
                    nodes = np.ascontiguousarray(nodes, dtype=np.float32)
                    keep.append(nodes)
                    pd.nodes = nodes.ctypes.data_as(_abi.c_float_p)
                mu = _phase_leaf_mueller(ph)
                if mu is not None:
                    for k, arr in enumerate(mu):
                        arr = np.ascontiguousarray(arr, dtype=np.float32)
                        keep.append(arr)
                        pd.mueller[k] = arr.ctypes.data_as(_abi.c_float_p)

        d.bsdf_type = self.bsdf_type()
        d.bsdf_params[:] = list(self.bsdf_params())
        if self.bsdf.type == "mqdiffuse":
            tab = self.bsdf.table
            keep.append(tab)
            d.bsdf_table = tab.ctypes.data_as(_abi.c_float_p)
            d.bsdf_table_res[:] = [tab.shape[2], tab.shape[1], tab.shape[0]]
        elif self.bsdf.type == "measured_mono":
            tab = self.bsdf.measured.table(self.bsdf.values["wavelength"])
            keep.append(tab)
            d.bsdf_table = tab.ctypes.data_as(_abi.c_float_p)
            d.bsdf_table_res[:] = [tab.size, 1, 1]
        d.emitter_direction[:] = list(self.emitter.direction)
        d.irradiance = self.emitter.children["irradiance"].values["value"]
        d.emitter_angular_diameter = getattr(self.emitter, "angular_diameter", 0.0)
        d.hide_emitters = int(bool(getattr(self.integrator, "hide_emitters", False)))
        it = self.integrator
        d.integrator = (
            {"volpathmis": _abi.INTEGRATOR_VOLPATHMIS,
             "piecewise_volpath": _abi.INTEGRATOR_PIECEWISE_VOLPATH}.get(it.kernel_type, _abi.INTEGRATOR_VOLPATH)
        )
        d.rr_depth = it.rr_depth
        d.max_depth = it.max_depth
        d.polarized = int(self.polarized)
        d.meridian_align = int(getattr(it, "meridian_align", False))

        sens = (_abi.SensorDesc * len(self.sensors))()
        for i, s in enumerate(self.sensors):
            sd = sens[i]
            sd.type = {
                "mdistant": _abi.SENSOR_MDISTANT,
                "hdistant": _abi.SENSOR_HDISTANT,
                "distantflux": _abi.SENSOR_DISTANTFLUX,
                "perspective": _abi.SENSOR_PERSPECTIVE,
                "mpdistant": _abi.SENSOR_MPDISTANT,
                "mradiancemeter": _abi.SENSOR_MRADIANCEMETER,
            }[s.type]
            if s.type == "mradiancemeter":
                org = np.ascontiguousarray(s.origins, dtype=np.float64)
                keep.append(org)
                sd.origins = org.ctypes.data_as(_abi.c_double_p)
                sd.in_medium = int(s.in_medium)
            if s.type == "perspective":
                sd.x_fov_deg, sd.near_clip, sd.far_clip = s.values["x_fov"], s.values["near_clip"], s.values["far_clip"]
                sd.in_medium = int(s.in_medium)
            sd.width, sd.height = s.film().width, s.film().height
            if s.type in ("mdistant", "mradiancemeter"):
                dirs = np.ascontiguousarray(s.directions, dtype=np.float64)
                keep.append(dirs)
                sd.n_directions = dirs.shape[0]
                sd.directions = dirs.ctypes.data_as(_abi.c_double_p)
            sd.to_world[:] = list(s.to_world.reshape(-1))
            sd.target_type = s.target_type
            sd.target[:] = list(s.target_point)
            sd.target_to_world[:] = list(s.target_to_world.reshape(-1))
            sd.ray_offset = s.ray_offset
        keep.append(sens)
        d.n_sensors = len(self.sensors)
        d.sensors = C.cast(sens, C.POINTER(_abi.SensorDesc))
        if self.patch_bsdf is not None:
            d.has_patch = 1
            d.patch_bsdf_type = self.bsdf_type(self.patch_bsdf)
            d.patch_bsdf_params[:] = list(self.bsdf_params(self.patch_bsdf))
            d.patch_rect[:] = list(self.patch_rect)
        if self.instances:
            groups = (_abi.LeafGroupDesc * len(self.leaf_groups))()
            for i, g in enumerate(self.leaf_groups):
                disks = np.ascontiguousarray(g.disks, dtype=np.float32)
                keep.append(disks)
                groups[i].n_disks = disks.shape[0]
                groups[i].disks = disks.ctypes.data_as(_abi.c_float_p)
                groups[i].reflectance, groups[i].transmittance = self.leaf_bsdf_params(i)
                if g.triangles.shape[0]:
                    tri = np.ascontiguousarray(g.triangles, dtype=np.float32)
                    tid = np.ascontiguousarray(g.triangle_bsdf, dtype=np.int32)
                    mb = np.ascontiguousarray(self.mesh_bsdf_params(i), dtype=np.float32)
                    keep += [tri, tid, mb]
                    groups[i].n_triangles, groups[i].n_mesh_bsdfs = tri.shape[0], mb.shape[0]
                    groups[i].triangles = tri.ctypes.data_as(_abi.c_float_p)
                    groups[i].triangle_bsdf = tid.ctypes.data_as(C.POINTER(C.c_int32))
                    groups[i].mesh_bsdfs = mb.ctypes.data_as(_abi.c_float_p)
                cyl = np.ascontiguousarray(g.cylinders, dtype=np.float32)
                tdk = np.ascontiguousarray(g.trunk_disks, dtype=np.float32)
                keep += [cyl, tdk]
                groups[i].n_cylinders, groups[i].n_trunk_disks = cyl.shape[0], tdk.shape[0]
                if cyl.shape[0]:
                    groups[i].cylinders = cyl.ctypes.data_as(_abi.c_float_p)
                if tdk.shape[0]:
                    groups[i].trunk_disks = tdk.ctypes.data_as(_abi.c_float_p)
                groups[i].trunk_reflectance = self.trunk_reflectance(i)
            inst_g = np.ascontiguousarray([gi for gi, _ in self.instances], dtype=np.int32)
            inst_o = np.ascontiguousarray([off for _, off in self.instances], dtype=np.float64)
            keep += [groups, inst_g, inst_o]
            d.n_leaf_groups, d.n_instances = len(self.leaf_groups), len(self.instances)
            d.leaf_groups = C.cast(groups, C.POINTER(_abi.LeafGroupDesc))
            d.instance_group = inst_g.ctypes.data_as(C.POINTER(C.c_int32))
            d.instance_offset = inst_o.ctypes.data_as(_abi.c_double_p)
        d._keepalive = keep  # the buffers live as long as the descriptor that points to them
        self._keepalive = keep
        return d


class _RayleighLeaf:
    """One constant-depolarization half of a Rayleigh leaf whose depolarization is a per-layer profile."""

    def __init__(self, type_: str, rho: float):
        self.type, self.rho = type_, rho


def _phase_leaf_mueller(ph: PhaseFunction):
    """m12, m22, m33, m34, m44 arrays of a tabphase_polarized leaf (else None)."""
    if ph.type != "tabphase_polarized":
        return None
    return [ph.values[k] for k in ("m12", "m22", "m33", "m34", "m44")]


def _phase_leaf_desc(ph: PhaseFunction):
    params = [0.0, 0.0, 0.0, 0.0]
    if ph.type == "isotropic":
        return _abi.PHASE_ISOTROPIC, params, None, None
    if ph.type == "tabphase_polarized":
        return _abi.PHASE_TABULATED_POLARIZED, params, ph.values["m11"], ph.values["nodes"]
    if ph.type in ("rayleigh", "rayleigh_polarized"):
        if isinstance(ph, _RayleighLeaf):
            params[0] = ph.rho
        else:
            dep = ph.children["depolarization"].layer_values()
            assert dep.size == 1  # (gridded profiles were split by phase_leaves)
            params[0] = float(dep.flat[0])
        if params[0] >= 1.0:
            raise RuntimeError("Depolarization factor must be in [0, 1[")
        ty = _abi.PHASE_RAYLEIGH_POLARIZED if ph.type == "rayleigh_polarized" else _abi.PHASE_RAYLEIGH
        return ty, params, None, None
    if ph.type == "hg":
        g = ph.values["g"]
        if not (-1.0 < g < 1.0):
            raise RuntimeError("The asymmetry parameter must lie in the interval (-1, 1)!")
        params[0] = g
        return _abi.PHASE_HG, params, None, None
    if ph.type == "tabphase":
        return _abi.PHASE_TABULATED, params, ph.values["values"], None
    if ph.type == "tabphase_irregular":
        return _abi.PHASE_TABULATED_IRREGULAR, params, ph.values["values"], ph.values["nodes"]
    raise RuntimeError(f"unsupported phase function '{ph.type}'")


def _shape_bbox(s: Shape):
    if s.type == "sphere":
        return s.center - s.radius, s.center + s.radius
    if s.type == "cube":
        corners = np.array(
            [[x, y, z, 1.0] for x in (-1, 1) for y in (-1, 1) for z in (-1, 1)], dtype=np.float64
        )
    else:  # rectangle / arectangle: [-1,1]^2 in the local XY plane
        corners = np.array([[x, y, 0.0, 1.0] for x in (-1, 1) for y in (-1, 1)], dtype=np.float64)
    w = (s.to_world @ corners.T).T[:, :3]
    return w.min(axis=0), w.max(axis=0)
