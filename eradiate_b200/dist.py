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

from .kernel._render import _device_scene, develop


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
