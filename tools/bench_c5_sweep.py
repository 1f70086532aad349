#!/usr/bin/env python
"""
C5-style sweep (BASELINE.json configs[4]): polarized ocean + molecular + polarized aerosol, one context per
band of a 400-1000 nm sweep, spp = 2^20 per band, bands dealt round-robin to the ranks (no reduction) and
queued through the pipelined batch entry points on each rank.

    python tools/bench_c5_sweep.py [--bands 61]                      # 1 GPU
    torchrun --nproc-per-node G tools/bench_c5_sweep.py [--bands 61]  # G GPUs

Rank 0 prints one JSON line (bands/s, Mpaths/s, checksum of the Stokes film means).
"""
import argparse, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from eradiate_b200 import scenes
from eradiate_b200.dist import mi_render_sharded
from eradiate_b200.kernel import KernelContext, mi_load_dict, mi_traverse
from eradiate_b200.kernel._render import SeedState


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bands", type=int, default=61)
    ap.add_argument("--spp", type=int, default=1 << 20)
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", 1))
    rank = int(os.environ.get("RANK", 0))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    sc = mi_load_dict(scenes.config_c5(spp=args.spp))
    mi_scene = mi_traverse(sc, scenes.spectral_update_map_c5(1200, True))
    ctxs = [KernelContext(w=w) for w in np.linspace(400.0, 1000.0, args.bands)]
    mi_render_sharded(mi_scene, ctxs[: 2 * world], spp=1 << 12, seed_state=SeedState(1))  # warm-up
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    res = mi_render_sharded(mi_scene, ctxs, spp=args.spp, seed_state=SeedState(0))
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    dt = time.perf_counter() - t0
    if rank == 0:
        s0 = np.array([res[c.si.as_hashable]["measure"].raw["sum_stokes"][0].sum() / args.spp for c in ctxs])
        s1 = np.array([res[c.si.as_hashable]["measure"].raw["sum_stokes"][1].sum() / args.spp for c in ctxs])
        print(json.dumps({"n_gpus": world, "bands": args.bands, "spp": args.spp, "seconds": round(dt, 4),
                          "bands_per_s": round(args.bands / dt, 1), "Mpaths_per_s": round(args.bands * args.spp / dt / 1e6, 1),
                          "I_400": float(s0[0]), "I_1000": float(s0[-1]), "Q_400": float(s1[0]), "checksum_I": float(s0.sum()),
                          "finite": bool(np.isfinite(s0).all())}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
