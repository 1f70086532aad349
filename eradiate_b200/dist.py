"""
Sample-sharded multi-GPU rendering (one process per GPU, torch.distributed / NCCL).

Every path is an independent Monte Carlo sample keyed by (seed, pixel, sample index),
so rank r of G renders the disjoint sample range [r*spp_rank, (r+1)*spp_rank) of EVERY
pixel (scene tables are < 64 KB and replicated) and the ranks exchange nothing until the
end, when ONE all-reduce sums the per-pixel accumulators (sum w*L, sum L, sum L^2:
3 x n_pixels float64 -- 768 B for mdistant-32).  PyTorch only provides the device buffer,
the stream and the NCCL collective; the path tracing itself is the CUDA library.
"""

from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

from .kernel._render import SeedState, _active_sensors, _device_scene, develop, get_seed_state


def is_distributed() -> bool:
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def shard_range(spp_total: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous sample range of ``rank``: (offset, count); counts differ by at most 1."""
    base, rem = divmod(int(spp_total), int(world))
    count = base + (1 if rank < rem else 0)
    offset = rank * base + min(rank, rem)
    return offset, count


class ShardedRenderer:
    """Holds the device accumulators and renders this rank's shard of a sensor."""

    def __init__(self, scene, device: int | None = None):
        if device is None:
            device = torch.cuda.current_device()
        self.device = device
        self.scene = scene
        self.dev = _device_scene(scene, device)
        self._accum: dict[int, torch.Tensor] = {}
        self._host: dict[int, torch.Tensor] = {}

    def accum(self, sensor: int) -> torch.Tensor:
        if sensor not in self._accum:
            npix = self.dev.lib.ertb_sensor_pixel_count(self.dev.handle, sensor)
            rows = 7 if self.dev.flat.polarized else 3  # + [S0 | S1 | S2 | S3] for polarized scenes
            self._accum[sensor] = torch.zeros(rows * npix, dtype=torch.float64, device=f"cuda:{self.device}")
            self._host[sensor] = torch.zeros(rows * npix, dtype=torch.float64).pin_memory()
        return self._accum[sensor]

    def launch(self, sensor: int, seed: int, spp: int, sample_offset: int = 0, stats: torch.Tensor | None = None):
        """Asynchronous: zero the accumulators and enqueue the render kernel on torch's current stream."""
        acc = self.accum(sensor)
        acc.zero_()
        stream = torch.cuda.current_stream(self.device).cuda_stream
        self.dev.render_device(sensor, seed, spp, sample_offset, acc.data_ptr(),
                               stats.data_ptr() if stats is not None else None, stream)
        return acc

    def render(self, sensor: int, seed: int, spp_total: int, to_host: bool = True):
        """
        Render ``spp_total`` samples per pixel over all ranks; returns the globally reduced
        (sum_wl, sum_l, sum_l2) as float64 numpy arrays (or the device tensor).
        """
        rank = dist.get_rank() if is_distributed() else 0
        world = dist.get_world_size() if is_distributed() else 1
        offset, count = shard_range(spp_total, rank, world)
        acc = self.accum(sensor)
        if count > 0:
            self.launch(sensor, seed, count, offset)
        else:
            acc.zero_()
        if world > 1:
            dist.all_reduce(acc, op=dist.ReduceOp.SUM)
        if not to_host:
            return acc
        host = self._host[sensor]
        host.copy_(acc, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        a = host.numpy().reshape(7 if self.dev.flat.polarized else 3, -1)
        self.last_stokes = a[3:7].copy() if self.dev.flat.polarized else None
        return a[0].copy(), a[1].copy(), a[2].copy()

    def render_bitmap(self, sensor: int, seed: int, spp_total: int):
        wl, l, l2 = self.render(sensor, seed, spp_total)
        return develop(self.scene, sensor, wl, l, l2, spp_total, stokes=self.last_stokes)


def reduce_host_accumulators(arrays, group=None):
    """All-reduce float64 host arrays (gloo): used by the CPU tests of the sharding logic."""
    out = []
    for a in arrays:
        t = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64).copy())
        if is_distributed():
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        out.append(t.numpy())
    return out


# ------------------------------------------------------------------------------
#      context ("band") sharding of the spectral loop: BASELINE config C5
# ------------------------------------------------------------------------------


def context_shard(n_items: int, rank: int, world: int) -> list[int]:
    """Items of the contexts x sensors loop dealt round-robin: rank r takes r, r+G, r+2G, ..."""
    return list(range(int(rank), int(n_items), int(world)))


def _render_items_gpu(mi_scene, plan, mine, seeds, spps, offsets=None):
    """This rank's items through the pipelined batch entry points (one device per rank)."""
    # the device of THIS rank (torchrun: set_device(LOCAL_RANK)), not device 0 of the box
    dev = _device_scene(mi_scene.obj, torch.cuda.current_device())
    dev.batch_begin([plan[k][1] for k in mine])
    last_ctx = None
    for k in mine:
        ctx, i_sensor, _ = plan[k]
        if ctx is not last_ctx:
            mi_scene.parameters.update(mi_scene.umap_template.render(ctx))
            last_ctx = ctx
        dev.batch_push(i_sensor, seeds[k], spps[k], 0 if offsets is None else offsets[k])
    items, _, _ = dev.batch_end()
    return items


def mi_render_sharded(mi_scene, ctxs, spp: int = 0, seed_state: SeedState | None = None,
                      render_items=_render_items_gpu, group=None, shard: str = "contexts") -> dict:
    """
    ``mi_render`` over all ranks; every rank returns the complete
    ``{ctx.si.as_hashable: {sensor_id: Bitmap}}``. Every rank draws ALL seeds in the global loop
    order, so item k gets the same seed whatever the world size.

    ``shard="contexts"`` (BASELINE C5, SURVEY 8e "band sharding"): the (context, sensor) items are
    dealt round-robin to the ranks, each rendered at full spp by one rank; the films are gathered,
    no reduction is involved, and the result does not depend on the number of GPUs at all.

    ``shard="samples"`` (BASELINE C3): every rank renders a disjoint range of the samples of EVERY
    item (same seed, sample offsets) and ONE all-reduce sums the accumulators of all items.
    """
    if shard not in ("contexts", "samples"):
        raise ValueError("shard must be 'contexts' or 'samples'")
    if seed_state is None:
        seed_state = get_seed_state()
    rank = dist.get_rank(group) if is_distributed() else 0
    world = dist.get_world_size(group) if is_distributed() else 1
    plan = [(ctx, i, s) for ctx in ctxs for i, s in _active_sensors(mi_scene, ctx)]
    seeds = [int(np.asarray(seed_state.next()).squeeze()) & 0xFFFFFFFFFFFFFFFF for _ in plan]
    spps = [int(spp) if spp > 0 else s.sampler().sample_count for _, _, s in plan]
    if shard == "samples":
        ranges = [shard_range(n, rank, world) for n in spps]
        mine = [k for k in range(len(plan)) if ranges[k][1] > 0]
        local = render_items(mi_scene, plan, mine, seeds, {k: ranges[k][1] for k in mine},
                             {k: ranges[k][0] for k in mine}) if mine else []
        shapes = [(7 if mi_scene.obj.flat.polarized else 3, s.film().width * s.film().height) for _, _, s in plan]
        flat = np.zeros(sum(r * n for r, n in shapes), dtype=np.float64)
        pos = np.cumsum([0] + [r * n for r, n in shapes])
        for k, a in zip(mine, local):
            flat[pos[k]:pos[k + 1]] = np.asarray(a, dtype=np.float64).ravel()
        if world > 1:
            t = torch.from_numpy(flat)
            if dist.get_backend(group) == "nccl":
                t = t.cuda()
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
            flat = t.cpu().numpy()
        payload = {k: flat[pos[k]:pos[k + 1]].reshape(shapes[k]) for k in range(len(plan))}
    else:
        mine = context_shard(len(plan), rank, world)
        local = render_items(mi_scene, plan, mine, seeds, spps) if mine else []
        payload = {k: np.asarray(a) for k, a in zip(mine, local)}
        if world > 1:
            gathered = [None] * world
            dist.all_gather_object(gathered, payload, group=group)
            payload = {k: a for part in gathered for k, a in part.items()}
    results: dict = {}
    for k, (ctx, i_sensor, mi_sensor) in enumerate(plan):
        a = payload[k]
        bmp = develop(mi_scene.obj, i_sensor, a[0], a[1], a[2], spps[k],
                      stokes=a[3:7] if a.shape[0] == 7 else None)
        results.setdefault(ctx.si.as_hashable, {})[mi_sensor.id()] = bmp
    return results
