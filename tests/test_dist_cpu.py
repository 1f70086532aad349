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


# --------------------------------------------------------------------------- band sharding
def _oracle_items(mi_scene, plan, mine, seeds, spps, offsets=None):
    """Stand-in for the GPU batch (test infrastructure): the CPU oracle renders this rank's items."""
    from oracle import oracle

    out = []
    for k in mine:
        ctx, i_sensor, _ = plan[k]
        mi_scene.parameters.update(mi_scene.umap_template.render(ctx))
        wl, l, l2, _ = oracle.render(mi_scene.obj.flat.build_desc(), i_sensor, seeds[k], spps[k],
                                     sample_offset=0 if offsets is None else offsets[k], n_threads=2)
        out.append(np.stack([wl, l, l2]))
    return out


def _band_scene():
    from eradiate_b200 import scenes
    from eradiate_b200.kernel import KernelContext, mi_load_dict, mi_traverse

    kdict = scenes.config_c2(spp=64, n_vza=3)
    kdict["measure_2"] = dict(kdict["measure"], id="measure_2")
    mi_scene = mi_traverse(mi_load_dict(kdict), scenes.spectral_update_map(1200, spherical=True))
    ctxs = [KernelContext(w=w) for w in (440.0, 550.0, 670.0, 865.0, 1020.0)]
    return mi_scene, ctxs


SPP_BAND = 65  # odd on purpose: ragged sample shards


def _band_worker(rank, world, port, out, shard):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from eradiate_b200.dist import mi_render_sharded
    from eradiate_b200.kernel._render import SeedState

    mi_scene, ctxs = _band_scene()
    res = mi_render_sharded(mi_scene, ctxs, spp=SPP_BAND, seed_state=SeedState(3), render_items=_oracle_items, shard=shard)
    if rank == 1:  # every rank holds the complete result
        np.save(out, np.stack([res[c.si.as_hashable][s].raw["sum_l"] for c in ctxs for s in ("measure", "measure_2")]))
    dist.barrier()
    dist.destroy_process_group()


def test_context_shard_is_a_partition():
    from eradiate_b200.dist import context_shard

    for n, world in ((10, 4), (3, 8), (960, 8), (1, 1)):
        parts = [context_shard(n, r, world) for r in range(world)]
        assert sorted(k for p in parts for k in p) == list(range(n))
        assert max(map(len, parts)) - min(map(len, parts)) <= 1


@pytest.mark.timeout(300)
@pytest.mark.parametrize("shard", ["contexts", "samples"])
def test_two_rank_gloo_sharded_mi_render_equals_single(tmp_path, oracle, shard):
    from eradiate_b200.dist import mi_render_sharded
    from eradiate_b200.kernel._render import SeedState

    out = str(tmp_path / "bands.npy")
    port = 31500 + (os.getpid() % 2000) + (7 if shard == "samples" else 0)
    mp.spawn(_band_worker, args=(2, port, out, shard), nprocs=2, join=True)
    got = np.load(out)
    mi_scene, ctxs = _band_scene()
    res = mi_render_sharded(mi_scene, ctxs, spp=SPP_BAND, seed_state=SeedState(3), render_items=_oracle_items)
    assert list(res.keys()) == [c.si.as_hashable for c in ctxs]
    want = np.stack([res[c.si.as_hashable][s].raw["sum_l"] for c in ctxs for s in ("measure", "measure_2")])
    assert got.shape == want.shape and np.allclose(got, want, rtol=1e-12)
    assert not np.allclose(want[0], want[1])  # the two sensors of a context got different seeds


def test_reference_arm_under_torchrun_prints_one_line_and_uses_all_cores():
    """bench.py --impl reference launched as the driver launches it for N > 1: rank 0 alone works and prints the
    one JSON line, the other rank exits 0 silently, and the oracle runs on every host core although torchrun
    exports OMP_NUM_THREADS=1 to its workers."""
    import json
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29577", os.path.join(root, "bench.py"),
           "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "3"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["gpu_launches"] == 0
    cores = os.cpu_count() or 1
    from oracle import ref

    kind = "reference" if ref.available() else "port"  # the reference itself when oracle/_ref was built
    assert d["cpu_baseline"]["cores"] == cores and d["cpu_baseline"]["kind"] == kind
    if cores >= 4:  # one thread manages ~0.09 Mpaths/s (reference, scalar_mono) / ~0.7 Mpaths/s (C port)
        assert d["value"] > (0.2 if kind == "reference" else 1.2), d["value"]
