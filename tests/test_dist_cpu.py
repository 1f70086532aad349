"""
N > 1 path on CPU: world_size-2 gloo processes shard the samples of every pixel, render
their shard (with the CPU oracle standing in for the GPU kernel -- test infrastructure),
all-reduce the accumulators, and must reproduce the single-process result exactly
(paths are keyed by (seed, pixel, sample index), so shards are disjoint subsets of the
same path set).
"""

import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, spp_total, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from eradiate_b200 import scenes
    from eradiate_b200.dist import reduce_host_accumulators, shard_range
    from eradiate_b200.kernel import mi_load_dict
    from oracle import oracle

    sc = mi_load_dict(scenes.config_c2(spp=spp_total, n_vza=4))
    desc = sc.flat.build_desc()
    off, cnt = shard_range(spp_total, rank, world)
    wl, l, l2, st = oracle.render(desc, 0, 42, cnt, sample_offset=off, n_threads=2)
    wl, l, l2 = reduce_host_accumulators([wl, l, l2])
    n = torch.tensor([st["n_paths"]], dtype=torch.int64)
    dist.all_reduce(n)
    if rank == 0:
        np.save(out, np.stack([wl, l, l2, np.full_like(wl, float(n[0]))]))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_range_partitions_samples():
    from eradiate_b200.dist import shard_range

    for spp, world in ((1 << 20, 8), (1000, 3), (5, 8), (7, 2)):
        ranges = [shard_range(spp, r, world) for r in range(world)]
        assert sum(c for _, c in ranges) == spp
        pos = 0
        for off, cnt in ranges:
            assert off == pos
            pos += cnt
        assert max(c for _, c in ranges) - min(c for _, c in ranges) <= 1


@pytest.mark.timeout(300)
def test_two_rank_gloo_sharded_render_equals_single(tmp_path, oracle):
    from eradiate_b200 import scenes
    from eradiate_b200.kernel import mi_load_dict

    spp = 3001  # odd on purpose: ragged shards
    out = str(tmp_path / "acc.npy")
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, spp, out), nprocs=2, join=True)
    got = np.load(out)
    sc = mi_load_dict(scenes.config_c2(spp=spp, n_vza=4))
    wl, l, l2, st = oracle.render(sc.flat.build_desc(), 0, 42, spp)
    assert got[3, 0] == st["n_paths"] == 4 * spp
    assert np.allclose(got[0], wl, rtol=1e-12)
    assert np.allclose(got[1], l, rtol=1e-12)
    assert np.allclose(got[2], l2, rtol=1e-12)
