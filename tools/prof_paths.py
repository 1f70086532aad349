#!/usr/bin/env python
"""Developer probe: kernel time of the C2 render (lean pool-kernel instance) after different preambles."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from eradiate_b200 import scenes
from eradiate_b200.dist import ShardedRenderer
from eradiate_b200.kernel import mi_load_dict
SPP = 1 << 20
which = sys.argv[1] if len(sys.argv) > 1 else "none"
scene = mi_load_dict(scenes.config_c2(spp=SPP, n_vza=32))
R = ShardedRenderer(scene, 0)
dev = R.dev
acc = R.accum(0)
stats = torch.zeros(8, dtype=torch.int64, device="cuda")
seed = 20261017
def timed(label, n=20):
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    torch.cuda.synchronize()
    for i, (a, b) in enumerate(evs):
        acc.zero_(); a.record()
        dev.render_device(0, seed + 1000 + i, SPP, 0, acc.data_ptr(), None, torch.cuda.current_stream().cuda_stream)
        b.record()
    torch.cuda.synchronize()
    ts = [a.elapsed_time(b) for a, b in evs]
    print(f"[{which}] {label:40s} median {np.median(ts):.3f} ms  min {np.min(ts):.3f}", flush=True)
if which == "none":
    pass
if which in ("stats_launch", "all"):
    R.launch(0, seed, SPP, 0, stats=stats); torch.cuda.synchronize()
if which in ("tiny_render", "all"):
    dev.render(0, seed, 16)
if which in ("warmups", "all"):
    for i in range(5): R.launch(0, seed + i, SPP, sample_offset=0)
    torch.cuda.synchronize()
timed("lean, first timing")
timed("lean, second timing")
flush = torch.empty(512 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")
def timed2(label, n=20, do_flush=True, sync_each=False, st=None, small=False):
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    torch.cuda.synchronize()
    for i, (a, b) in enumerate(evs):
        if do_flush:
            (flush[:1 << 20] if small else flush).fill_(float(i))
        acc.zero_(); a.record()
        dev.render_device(0, seed + 1000 + i, SPP, 0, acc.data_ptr(), st, torch.cuda.current_stream().cuda_stream)
        b.record()
        if sync_each: torch.cuda.synchronize()
    torch.cuda.synchronize()
    ts = [a.elapsed_time(b) for a, b in evs]
    print(f"[{which}] {label:52s} median {np.median(ts):.3f} ms  min {np.min(ts):.3f}  max {np.max(ts):.3f}", flush=True)
timed2("lean, 512 MiB fill before, no host sync")
timed2("lean, 512 MiB fill before, host sync each")
timed2("lean, 4 MiB fill before, no host sync", small=True)
timed2("lean, no fill, no host sync", do_flush=False)
timed2("stats+COLL, 512 MiB fill before, no host sync", st=stats.data_ptr())
