"""
Drop-in replacements for ``eradiate.kernel.mi_load_dict / mi_traverse / mi_render``
(``src/eradiate/kernel/_render.py:186, :212, :379``) backed by the sm_100a CUDA
path tracer instead of Mitsuba.

Same names, argument meaning, return structure and error behaviour
(``RuntimeError`` on load failures, ``warnings.warn`` on unsuccessful parameter
lookups).  The CUDA library is mandatory: there is no CPU fallback.
"""

from __future__ import annotations

import ctypes as C
import logging
import re
import typing as t
import warnings

import numpy as np

from .. import _abi, _lib
from . import _scene
from ._bitmap import Bitmap
from ._kernel_dict import KernelSceneParameterMap, SceneParameter

logger = logging.getLogger(__name__)

# ------------------------------------------------------------------------------
#                                seed state
# ------------------------------------------------------------------------------


class SeedState:
    """``src/eradiate/rng.py``: numpy ``SeedSequence``-based seed generator."""

    def __init__(self, seed=None):
        self._seed = (
            seed if isinstance(seed, np.random.SeedSequence) else np.random.SeedSequence(seed)
        )

    def reset(self, seed=None):
        if seed is not None:
            self._seed = (
                seed if isinstance(seed, np.random.SeedSequence) else np.random.SeedSequence(seed)
            )
        else:
            self._seed = np.random.SeedSequence(entropy=self._seed.entropy)

    def next(self, n: int = 1) -> np.ndarray:
        return self._seed.spawn(1)[0].generate_state(n)


_root_seed_state = SeedState(0)  # src/eradiate/config/_defaults.py:50 rng_seed = 0


def get_seed_state() -> SeedState:
    return _root_seed_state


# ------------------------------------------------------------------------------
#                        device-side scene (C ABI handle)
# ------------------------------------------------------------------------------


class DeviceScene:
    """Owns one ``ertb_scene`` handle and keeps it in sync with the host scene graph."""

    def __init__(self, scene: _scene.Scene, device: int = 0):
        self.scene = scene
        self.device = device
        self.lib = _lib.load()
        self.flat = scene.flat
        self.desc = self.flat.build_desc()
        handle = C.c_void_p()
        _lib.check(self.lib.ertb_scene_create(C.byref(self.desc), device, C.byref(handle)))
        self.handle = handle
        self._node_list = None
        self._batch = None
        self._batch_open = False
        self._mark_clean()

    def _nodes(self):
        """Every node of the scene graph (the graph is fixed after loading: collected once)."""
        if self._node_list is None:
            out, stack, seen = [], [self.scene], set()
            while stack:
                n = stack.pop()
                if id(n) in seen:
                    continue
                seen.add(id(n))
                out.append(n)
                stack.extend(n.children.values())
            self._node_list = out
        return self._node_list

    @staticmethod
    def _subtree_dirty(root) -> bool:
        stack = [root]
        while stack:
            n = stack.pop()
            if n.dirty:
                return True
            stack.extend(n.children.values())
        return False

    def _mark_clean(self):
        for n in self._nodes():
            n.dirty = False
            n.dirty_names.clear()

    # (plugin kind, value name) pairs ``ertb_scene_update`` can refresh in place; everything else the
    # parameter table publishes (medium ``scale``, the ``nodes`` of irregular / polarized tabulated phase
    # functions, sensor and emitter geometry ...) is honoured by re-creating the device scene from the
    # updated graph -- slower, never silently stale (Mitsuba's ``parameters_changed`` accepts all of them).
    _IN_PLACE = {
        ("volume", "data"), ("volume", "value"), ("texture", "value"),
        ("phase", "values"), ("phase", "g"), ("phase", "depolarization"),
        ("phase", "m11"), ("phase", "m12"), ("phase", "m22"), ("phase", "m33"), ("phase", "m34"), ("phase", "m44"),
    }

    def _needs_rebuild(self) -> bool:
        for n in self._nodes():
            if n.dirty_names and getattr(n, "rebuild_on_change", False):  # e.g. the index of a selectbsdf
                return True
            for name in n.dirty_names:
                if n.plugin_kind == "bsdf":  # every BSDF value is re-read by bsdf_params()
                    continue
                if (n.plugin_kind, name) not in self._IN_PLACE:
                    return True
        return False

    def _rebuild(self) -> None:
        """Re-create the device scene from the (updated) host graph."""
        if getattr(self, "_batch", None) is not None and self._batch_open:
            raise RuntimeError("a parameter that needs a scene rebuild was updated inside a pipelined batch")
        self.scene._flat = None
        self.flat = self.scene.flat
        self.desc = self.flat.build_desc()
        handle = C.c_void_p()
        _lib.check(self.lib.ertb_scene_create(C.byref(self.desc), self.device, C.byref(handle)))
        self.lib.ertb_scene_destroy(self.handle)
        self.handle = handle
        self.rebuilds = getattr(self, "rebuilds", 0) + 1
        self._mark_clean()

    def sync(self) -> None:
        """Push updated parameters (``parameters_changed`` equivalent)."""
        if not any(n.dirty for n in self._nodes()):
            return
        if self._needs_rebuild():
            self._rebuild()
            return
        flat, lib, h = self.flat, self.lib, self.handle

        def push(param, index, arr):
            arr = np.ascontiguousarray(arr, dtype=np.float32).reshape(-1)
            _lib.check(
                lib.ertb_scene_update(h, param, index, arr.ctypes.data_as(_abi.c_float_p), arr.size)
            )

        # only what changed is flattened again and pushed (a spectral loop touches sigma_t, albedo, the blend
        # weights and the irradiance of every context: ~0.1 ms of host work instead of ~0.5 ms for everything)
        dirty = self._subtree_dirty
        if flat.medium is not None:
            n = flat.n_layers()
            if n != self.desc.n_layers:
                raise RuntimeError("the number of atmospheric layers cannot change after loading")
            m = flat.medium
            if dirty(m.children["sigma_t"]):
                push(_abi.PARAM_SIGMA_T, 0, flat._profile(m.children["sigma_t"], n, "sigma_t"))
            if dirty(m.children["albedo"]):
                push(_abi.PARAM_ALBEDO, 0, flat._profile(m.children["albedo"], n, "albedo"))
            if dirty(m.children["phase_function"]):
                leaves = flat.phase_leaves(n)
                if len(leaves) != self.desc.n_phase:
                    raise RuntimeError("the phase function tree cannot change after loading")
                push(_abi.PARAM_PHASE_WEIGHT, 0, np.stack([p for _, p in leaves]))
                for i, (ph, _) in enumerate(leaves):
                    _, params, values, _ = _scene._phase_leaf_desc(ph)
                    push(_abi.PARAM_PHASE_PARAMS, i, np.asarray(params))
                    if values is not None:
                        push(_abi.PARAM_PHASE_VALUES, i, values)
                    mu = _scene._phase_leaf_mueller(ph)
                    if mu is not None:
                        for k, arr in enumerate(mu):
                            push(_abi.PARAM_PHASE_MUELLER, 5 * i + k, arr)
        for i in range(len(flat.leaf_groups) if flat.instances else 0):
            g = flat.leaf_groups[i]
            if dirty(g):
                push(_abi.PARAM_LEAF_BSDF, i, flat.leaf_bsdf_params(i))
                if "trunk_bsdf" in g.children:
                    push(_abi.PARAM_TRUNK_BSDF, i, [flat.trunk_reflectance(i)])
                for k, rt in enumerate(flat.mesh_bsdf_params(i)):
                    push(_abi.PARAM_MESH_BSDF, (i << 16) | k, rt)
        if flat.patch_bsdf is not None and dirty(flat.patch_bsdf):
            push(_abi.PARAM_PATCH_BSDF_PARAMS, 0, flat.bsdf_params(flat.patch_bsdf))
        if dirty(flat.bsdf):
            push(_abi.PARAM_BSDF_PARAMS, 0, flat.bsdf_params())
        if dirty(flat.emitter):
            push(_abi.PARAM_IRRADIANCE, 0, [flat.emitter.children["irradiance"].values["value"]])
        self._mark_clean()

    def render(self, sensor: int, seed: int, spp: int, sample_offset: int = 0, with_stats: bool = True):
        """One ``ertb_render`` call. Returns (sum_wl, sum_l, sum_l2, stats).  ``with_stats=False`` runs the
        kernel instance without loop-trip counters (what ``mi_render`` uses: ~5 % faster) and returns None for them."""
        self.sync()
        npix = self.lib.ertb_sensor_pixel_count(self.handle, sensor)
        if npix <= 0:
            raise RuntimeError(f"invalid sensor index {sensor}")
        out = [np.zeros(npix, dtype=np.float64) for _ in range(3)]
        stats = _abi.RenderStats() if with_stats else None
        stats_ref = C.byref(stats) if with_stats else None
        if self.flat.polarized:
            stokes = np.zeros((4, npix), dtype=np.float64)
            _lib.check(
                self.lib.ertb_render_stokes(
                    self.handle, sensor, seed, spp, sample_offset,
                    *[o.ctypes.data_as(_abi.c_double_p) for o in out],
                    stokes.ctypes.data_as(_abi.c_double_p), stats_ref,
                )
            )
            self.last_stokes = stokes
            return out[0], out[1], out[2], stats
        self.last_stokes = None
        _lib.check(
            self.lib.ertb_render(
                self.handle, sensor, seed, spp, sample_offset,
                *[o.ctypes.data_as(_abi.c_double_p) for o in out], stats_ref,
            )
        )
        return out[0], out[1], out[2], stats

    def render_device(self, sensor, seed, spp, sample_offset, accum_ptr, stats_ptr=None, stream=None):
        """Asynchronous render into a caller-owned device buffer (see the header)."""
        self.sync()
        _lib.check(
            self.lib.ertb_render_device(
                self.handle, sensor, seed, spp, sample_offset,
                C.c_void_p(accum_ptr), C.c_void_p(stats_ptr or 0), C.c_void_p(stream or 0),
            )
        )

    # -- pipelined contexts x sensors loop (ertb_batch_*) -------------------
    def batch_begin(self, sensors: list[int], with_stats: bool = False) -> None:
        arr = (C.c_int * len(sensors))(*sensors)
        self.sync()  # (a pending rebuild must not happen inside the batch)
        _lib.check(self.lib.ertb_batch_begin(self.handle, len(sensors), arr, int(with_stats)))
        self._batch = (list(sensors), with_stats)
        self._batch_open = True

    def batch_push(self, sensor: int, seed: int, spp: int, sample_offset: int = 0) -> None:
        """Snapshot the current parameters and queue one render; does not wait for it."""
        self.sync()
        _lib.check(self.lib.ertb_batch_push(self.handle, sensor, seed, spp, sample_offset))

    def batch_end(self):
        """Wait for the queued renders. Returns ([accumulators[rows, npix] per item], [stats], ms)."""
        sensors, with_stats = self._batch
        self._batch_open = False
        rows = 7 if self.flat.polarized else 3
        npix = [self.lib.ertb_sensor_pixel_count(self.handle, i) for i in sensors]
        out = np.zeros(rows * sum(npix), dtype=np.float64)
        stats = (_abi.RenderStats * len(sensors))() if with_stats else None
        ms = C.c_double(0.0)
        _lib.check(
            self.lib.ertb_batch_end(
                self.handle, out.ctypes.data_as(_abi.c_double_p), out.size, stats, C.byref(ms)
            )
        )
        items, o = [], 0
        for n in npix:
            items.append(out[o:o + rows * n].reshape(rows, n))
            o += rows * n
        return items, (list(stats) if with_stats else [None] * len(sensors)), ms.value

    def close(self):
        if getattr(self, "handle", None):
            self.lib.ertb_scene_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _default_device() -> int:
    """Device of a scene created without an explicit one: the calling process's current CUDA device when
    torch is in use (one process per GPU: ``torch.cuda.set_device(LOCAL_RANK)``), else device 0."""
    import sys

    torch = sys.modules.get("torch")
    if torch is not None and torch.cuda.is_available() and torch.cuda.is_initialized():
        return int(torch.cuda.current_device())
    return 0


def _device_scene(scene: _scene.Scene, device: int | None = None) -> DeviceScene:
    dev = getattr(scene, "_device_scene", None)
    want = _default_device() if device is None else device
    if dev is None or (device is not None and dev.device != want):
        if dev is not None:
            dev.close()
        dev = DeviceScene(scene, want)
        scene._device_scene = dev
    return dev


# ------------------------------------------------------------------------------
#                               scene parameters
# ------------------------------------------------------------------------------


class SceneParameters:
    """``mitsuba.SceneParameters`` stand-in: dotted key -> (node, value name)."""

    def __init__(self, properties: dict, scene=None, aliases: dict | None = None):
        self.properties = properties
        self.scene = scene
        self.aliases = aliases or {}
        self.update_candidates: dict = {}

    def __contains__(self, key):
        return self.aliases.get(key, key) in self.properties

    def __getitem__(self, key):
        node, name = self.properties[self.aliases.get(key, key)]
        return node.values[name]

    def __setitem__(self, key, value):
        key = self.aliases.get(key, key)
        if key not in self.properties:
            raise KeyError(key)
        self.update_candidates[key] = value

    def __len__(self):
        return len(self.properties)

    def keys(self):
        return self.properties.keys()

    def items(self):
        return ((k, self[k]) for k in self.properties)

    def keep(self, keys) -> None:
        if isinstance(keys, str):
            regexps = [re.compile(keys).fullmatch]
            keys = [k for k in self.properties if any(r(k) for r in regexps)]
        keys = [self.aliases.get(k, k) for k in keys]
        self.properties = {k: v for k, v in self.properties.items() if k in set(keys)}

    def update(self, values: dict | None = None) -> list:
        if values is not None:
            for k, v in values.items():
                if v is SceneParameter.UNUSED:
                    continue
                if k not in self:
                    raise KeyError(f"scene parameter '{k}' not found")
                self[k] = v
        touched = []
        for key, value in self.update_candidates.items():
            node, name = self.properties[key]
            node.set_value(name, value)
            touched.append((node, {name}))
        self.update_candidates = {}
        return touched

    def __repr__(self):
        lines = "\n".join(f"  {k}" for k in self.properties)
        return f"SceneParameters[\n{lines}\n]"


class MitsubaObjectWrapper:
    """``_render.py:74-140``: scene + parameter table + update-map template."""

    def __init__(self, obj, parameters=None, umap_template=None):
        self.obj = obj
        self.parameters = parameters
        self.umap_template = umap_template

    def drop_parameters(self) -> None:
        if self.umap_template is not None:
            keys = []
            for k, v in self.umap_template.items():
                if getattr(v, "parameter_id", None) is not None:
                    keys.append(v.parameter_id)
                else:
                    keys.append(k)
            self.parameters.keep(keys)

    def __repr__(self):
        return "MitsubaObjectWrapper[obj=Scene[...], parameters=SceneParameters[...]]"


# ------------------------------------------------------------------------------
#                       mi_load_dict / mi_traverse / mi_render
# ------------------------------------------------------------------------------


def mi_load_dict(dict: dict, parallel: bool = True, optimize: bool = False) -> object:
    """
    Load a scene (or a single plugin) from a Mitsuba-style dictionary
    (``_render.py:186-209``).  Unsupported plugins and malformed scenes raise
    ``RuntimeError``.
    """
    return _scene.load_dict(dict)


def mi_traverse(obj, umap_template=None, name_id_override=None) -> MitsubaObjectWrapper:
    """
    Traverse the scene graph and return the parameter table
    (``_render.py:212-371``).  Parameter lookups registered in ``umap_template``
    through ``search`` callables are resolved during traversal; unsuccessful
    lookups emit a warning.
    """
    umap_template = (
        KernelSceneParameterMap(data=dict(getattr(umap_template, "data", umap_template)))
        if umap_template is not None
        else KernelSceneParameterMap()
    )
    lookups = {
        k: v
        for k, v in umap_template.items()
        if getattr(v, "parameter_id", None) is None and getattr(v, "search", None) is not None
    }
    if name_id_override is None or name_id_override is False:
        name_id_override = []
    if name_id_override is True:
        name_id_override = [r".*"]
    if not isinstance(name_id_override, list):
        name_id_override = [name_id_override]
    regexps = [re.compile(k).match for k in name_id_override]

    properties: dict = {}
    hierarchy: dict = {}
    prefixes: set = set()
    aliases: dict = {}

    class SceneTraversal:
        def __init__(self, node, parent=None, name=None, depth=0):
            node_id = node.id()
            if name_id_override and node_id:
                for r in regexps:
                    if r(node_id):
                        if node_id != name:
                            aliases[node_id] = name
                        name = node_id
                        break
            if name is not None:
                ctr, name_len = 1, len(name)
                while name in prefixes:
                    name = f"{name[:name_len]}_{ctr}"
                    ctr += 1
                prefixes.add(name)
            self.name, self.node, self.depth = name, node, depth
            hierarchy[id(node)] = (parent, depth)
            for key, uparam in list(lookups.items()):
                found = uparam.search(self.node, self.name)
                if found is not None:
                    uparam.parameter_id = found
                    del lookups[key]

        def put(self, name, value, flags=0, cpptype=None):
            if isinstance(value, _scene.Object):
                self.put_object(name, value, flags)
            else:
                self.put_value(name, value, flags, cpptype)

        def put_value(self, name, value, flags=0, cpptype=None):
            full = name if self.name is None else f"{self.name}.{name}"
            properties[full] = (self.node, name)

        def put_object(self, name, child, flags=0):
            if child is None or id(child) in hierarchy:
                return
            cb = SceneTraversal(
                child,
                parent=self.node,
                name=name if self.name is None else f"{self.name}.{name}",
                depth=self.depth + 1,
            )
            child.traverse(cb)

    cb = SceneTraversal(obj)
    obj.traverse(cb)

    if lookups:
        warnings.warn(
            "There were unsuccessful Mitsuba scene parameter lookups: " f"{list(lookups.keys())}"
        )
    return MitsubaObjectWrapper(
        obj=obj,
        parameters=SceneParameters(properties, obj, aliases),
        umap_template=umap_template,
    )


def render(scene, sensor: int = 0, seed: int = 0, spp: int = 0, device: int | None = None, stats: bool = True):
    """
    ``mitsuba.render(scene, sensor=i, seed=seed, spp=spp)`` for scalar variants
    (``MI/src/python/python/util.py:511-519``): renders, develops the film and
    returns the bitmap (also available from ``sensor.film().bitmap()``).
    """
    if not isinstance(scene, _scene.Scene):
        raise RuntimeError("render(): expected a scene loaded with mi_load_dict")
    sensors = scene.sensors()
    if isinstance(sensor, int):
        if not 0 <= sensor < len(sensors):
            raise RuntimeError(f"render(): sensor index {sensor} out of range")
        i_sensor = sensor
    else:
        i_sensor = sensors.index(sensor)
    s = sensors[i_sensor]
    if spp <= 0:
        spp = s.sampler().sample_count
    dev = _device_scene(scene, device)
    sum_wl, sum_l, sum_l2, st = dev.render(i_sensor, int(seed) & 0xFFFFFFFFFFFFFFFF, int(spp), with_stats=stats)
    bmp = develop(scene, i_sensor, sum_wl, sum_l, sum_l2, spp, stokes=dev.last_stokes)
    bmp.stats = st.as_dict() if st is not None else None
    s.film()._bitmap = bmp
    return bmp


def develop(scene, i_sensor: int, sum_wl, sum_l, sum_l2, spp: int, stokes=None) -> Bitmap:
    """``HDRFilm::develop`` (``hdrfilm.cpp:304-405``): sums / weight, channel naming."""
    s = scene.sensors()[i_sensor]
    film = s.film()
    h, w = film.height, film.width
    inv = 1.0 / float(spp)
    chans = [np.asarray(sum_wl).reshape(h, w) * inv]
    names = ["Y"]
    if getattr(scene.integrator(), "stokes", False):
        # stokes.cpp:190-196: AOVs S0.R, S0.G, S0.B, S1.R ... come first
        st = np.zeros((4, h * w)) if stokes is None else np.asarray(stokes)
        for k in range(4):
            c = st[k].reshape(h, w) * inv
            chans += [c, c, c]
            names += [f"S{k}.R", f"S{k}.G", f"S{k}.B"]
    if scene.integrator().moment:
        m1 = np.asarray(sum_l).reshape(h, w) * inv
        m2 = np.asarray(sum_l2).reshape(h, w) * inv
        chans += [m1, m1, m1, m2, m2, m2]
        names += ["nested.X", "nested.Y", "nested.Z", "m2_nested.X", "m2_nested.Y", "m2_nested.Z"]
    data = np.stack(chans, axis=-1).astype(np.float32)
    raw = {
        "sum_wl": np.asarray(sum_wl, dtype=np.float64).reshape(h, w).copy(),
        "sum_l": np.asarray(sum_l, dtype=np.float64).reshape(h, w).copy(),
        "sum_l2": np.asarray(sum_l2, dtype=np.float64).reshape(h, w).copy(),
        "spp": int(spp),
    }
    if stokes is not None:
        raw["sum_stokes"] = np.asarray(stokes, dtype=np.float64).reshape(4, h, w).copy()
    return Bitmap(data, None, names, raw=raw)


def _active_sensors(mi_scene, ctx):
    sensors = mi_scene.obj.sensors()
    active_sensors = getattr(ctx, "active_sensors", None)
    if active_sensors is None:
        return list(enumerate(sensors))
    return [(i, sensors[i]) for i in active_sensors]


def _mi_render_pipelined(mi_scene, ctxs, spp, seed_state) -> dict:
    """
    Same loop, same seeds, same results as the sequential one below, but the renders are
    queued through ``ertb_batch_*``: the parameter update of context i+1 (host work + table
    upload) overlaps the render of context i and all films come back with one copy.
    """
    scene = mi_scene.obj
    plan = [(ctx, _active_sensors(mi_scene, ctx)) for ctx in ctxs]
    dev = _device_scene(scene)
    dev.batch_begin([i for _, act in plan for i, _ in act])
    spps = []
    for ctx, act in plan:
        mi_scene.parameters.update(mi_scene.umap_template.render(ctx))
        for i_sensor, mi_sensor in act:
            seed = int(np.asarray(seed_state.next()).squeeze())
            n = int(spp) if spp > 0 else mi_sensor.sampler().sample_count
            dev.batch_push(i_sensor, seed & 0xFFFFFFFFFFFFFFFF, n)
            spps.append(n)
    items, _, _ = dev.batch_end()
    results: dict = {}
    k = 0
    for ctx, act in plan:
        for i_sensor, mi_sensor in act:
            a = items[k]
            bmp = develop(scene, i_sensor, a[0], a[1], a[2], spps[k], stokes=a[3:7] if a.shape[0] == 7 else None)
            mi_sensor.film()._bitmap = bmp
            results.setdefault(ctx.si.as_hashable, {})[mi_sensor.id()] = bmp
            k += 1
    return results


def mi_render(
    mi_scene: MitsubaObjectWrapper,
    ctxs: list,
    spp: int = 0,
    seed_state: SeedState | None = None,
    pipelined: bool = True,
) -> dict[t.Any, dict[str, Bitmap]]:
    """
    Render the scene for every context and active sensor (``_render.py:379-470``).
    Returns ``{ctx.si.as_hashable: {sensor_id: Bitmap}}``.

    ``pipelined`` (extension, default on): queue the (context, sensor) renders through the
    asynchronous batch entry points instead of synchronising after each of them; the seeds
    drawn and the estimates returned are the same.
    """
    if seed_state is None:
        logger.debug("Using default RNG seed generator")
        seed_state = get_seed_state()

    if pipelined and sum(len(_active_sensors(mi_scene, c)) for c in ctxs) > 1:
        return _mi_render_pipelined(mi_scene, ctxs, spp, seed_state)

    results: dict = {}
    for ctx in ctxs:
        logger.debug("Updating scene parameters")
        mi_scene.parameters.update(mi_scene.umap_template.render(ctx))

        active_sensors = getattr(ctx, "active_sensors", None)
        sensors = mi_scene.obj.sensors()
        if active_sensors is None:
            mi_sensors = list(enumerate(sensors))
        else:
            mi_sensors = [(i, sensors[i]) for i in active_sensors]

        for i_sensor, mi_sensor in mi_sensors:
            seed = int(np.asarray(seed_state.next()).squeeze())
            logger.debug('Running kernel for sensor "%s" with seed value %s', mi_sensor.id(), seed)
            render(mi_scene.obj, sensor=i_sensor, seed=seed, spp=spp, stats=False)
            siah = ctx.si.as_hashable
            results.setdefault(siah, {})[mi_sensor.id()] = Bitmap(mi_sensor.film().bitmap())
    return results
