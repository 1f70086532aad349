#!/usr/bin/env python
"""
bench.py -- headline benchmark: Mpaths/s on the TOA-BRF AFGL1986+RPV scene
(BASELINE.json configs[1] "C2": AFGL-1986-shaped molecular atmosphere, RPV surface,
mono 550 nm, spherical shell, mdistant 32 VZA, spp = 2^20 per pixel per GPU).

    python bench.py --gpus N --steps K --warmup W          # CUDA path (this repo)
    python bench.py --impl reference --gpus N ...          # CPU arm: oracle port, all host cores

A "step" is one pass of the hot path: one render of the whole film (32 x 2^20 = 33.5 M
paths per GPU).  `value` is whole-job throughput with the scene tables resident in HBM
(device accumulators, CUDA-event timing on the launch stream, max over ranks, one NCCL
all-reduce of the accumulators inside the timed step when N > 1).  `e2e` is the same
metric through the public host API (parameter update from host memory -> render ->
accumulators copied back to the host).  One JSON line is printed by rank 0.
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "Mpaths/s TOA-BRF AFGL1986+RPV"
UNIT = "Mpaths/s"
SPP = 1 << 20
N_VZA = 32
# --config c3 (not the default, not the headline): BASELINE.json configs[2], the one named for
# sample sharding over 1 -> 8 GPUs, at a per-GPU size that keeps a step around a quarter of a second
C3_SPP = 1 << 19
C3_RES = 32
C3_WORKLOAD = ("C3: AFGL1986-shaped molecular atmosphere + aerosol layer (tabphase), spherical shell, "
               "hdistant 32x32, spp=2^19 per pixel per GPU, volpath (banded majorant)")
RECORD_BYTES = 64  # mono path record S = 64 B, one read + one write per loop trip
WORKLOAD = ("C2: AFGL1986-shaped molecular atmosphere (1200 layers) + RPV surface, mono 550 nm, "
            "spherical shell, mdistant 32 VZA, spp=2^20 per pixel per GPU, volpath (global majorant)")


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def measured_peak_gbs():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic_bytes():
    """dram bytes per launch of the render kernel from the committed ncu capture, or None."""
    path = os.path.join(ROOT, "profiles", "render_kernel_traffic.json")
    try:
        return float(json.load(open(path))["dram_bytes_per_launch"])
    except Exception:
        return None


def ncu_issue(kernel_ms, sm_mhz, n_sm):
    """Instruction-issue view of the same kernel (the bound DESIGN.md names): warp instructions per launch from
    the committed ncu capture / the kernel time measured live, against n_sm x 4 schedulers x 1 inst/clk."""
    path = os.path.join(ROOT, "profiles", "render_kernel_traffic.json")
    try:
        d = json.load(open(path))
        inst, lanes = float(d["warp_inst_per_launch"]), float(d["thread_inst_per_warp_inst"])
        achieved = inst / (kernel_ms * 1e-3) / 1e9
        peak = n_sm * 4 * float(sm_mhz) * 1e6 / 1e9
        return {"achieved": achieved, "peak": peak, "unit": "G warp-inst/s", "frac": achieved / peak,
                "active_lanes_per_inst": lanes,
                "note": "instruction count of one C2 launch from profiles/r01x (ncu), time measured live"}
    except Exception:
        return None


class ClockSampler(threading.Thread):
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons = index, [], set()
        self.max_mhz, self._stop_evt, self.proc = None, threading.Event(), None

    def run(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-i", str(self.index), "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                if self._stop_evt.is_set():
                    break
                f = [x.strip() for x in line.split(",")]
                if len(f) < 6:
                    continue
                try:
                    self.samples.append(float(f[0]))
                    self.max_mhz = float(f[1])
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                    "sw_power_cap"), f[2:6]):
                    if v.lower().startswith("active"):
                        self.reasons.add(name)
        except FileNotFoundError:
            pass

    def stop(self):
        self._stop_evt.set()
        if self.proc is not None:
            self.proc.terminate()
        self.join(timeout=2)
        return {
            "sm_mhz": float(np.median(self.samples)) if self.samples else None,
            "sm_max_mhz": self.max_mhz,
            "reasons": sorted(self.reasons),
            "samples": len(self.samples),
        }


# ------------------------------------------------------------------------------
#                       CPU arm: the oracle port on the host cores
# ------------------------------------------------------------------------------


def time_cpu_oracle(spp: int, repeats: int = 1, n_threads: int = 0):
    """Returns (Mpaths/s, seconds per call, K-bar) of the oracle on a bounded C2 sample."""
    from eradiate_b200 import scenes
    from eradiate_b200.kernel import mi_load_dict
    from oracle import oracle

    oracle.build()
    scene = mi_load_dict(scenes.config_c2(spp=spp, n_vza=N_VZA))
    desc = scene.flat.build_desc()
    oracle.render(desc, 0, 1, 64, n_threads=n_threads)  # warm the thread pool
    best = float("inf")
    kbar = 0.0
    for i in range(repeats):
        t0 = time.perf_counter()
        _, _, _, st = oracle.render(desc, 0, 100 + i, spp, n_threads=n_threads)
        best = min(best, time.perf_counter() - t0)
        kbar = (st["trips_main"] + st["trips_nee"]) / st["n_paths"]
    return N_VZA * spp / best / 1e6, best, kbar


def run_reference(args):
    rank = env_int("RANK", 0)
    if rank != 0:
        return  # other ranks exit 0 without work
    cores = os.cpu_count() or 1
    spp = 1 << 18  # bounded sample: 32 x 2^18 = 8.4 M paths per step (~1 s on 16 cores)
    # n_threads is passed explicitly: torchrun exports OMP_NUM_THREADS=1 to every rank
    for _ in range(args.warmup):
        time_cpu_oracle(1 << 12, n_threads=cores)
    times = []
    for _ in range(args.steps):
        mp, sec, _ = time_cpu_oracle(spp, n_threads=cores)
        times.append(sec)
    ms = 1e3 * float(np.mean(times))
    value = N_VZA * spp / (ms * 1e-3) / 1e6
    sample = (f"{args.steps} steps x (32 pixels x spp=2^18 = {N_VZA * spp} paths) of the C2 workload, "
              f"OpenMP on all {cores} host cores")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": sample,
                   "note": "reference's own implementation cannot be built/imported here (Mitsuba+Dr.Jit need "
                           "cmake + generated headers); this arm times the CPU oracle port of the same algorithm"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------
#                                   CUDA arm
# ------------------------------------------------------------------------------


def run_cuda(args):
    import torch
    import torch.distributed as dist

    import __graft_entry__ as g

    world = env_int("WORLD_SIZE", 1)
    rank = env_int("RANK", 0)
    local_rank = env_int("LOCAL_RANK", 0)
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (there is no CPU fallback); "
                           "use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL prints its version banner on stdout when it initialises (the GPU boxes export NCCL_DEBUG=VERSION),
        # in front of the one JSON line of the contract: point fd 1 at stderr while the communicator comes up
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
            warm = torch.zeros(1, device=f"cuda:{local_rank}")
            dist.all_reduce(warm)
            torch.cuda.synchronize()
        finally:
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)
    if rank == 0:
        g.build_cuda()
    if world > 1:
        dist.barrier()

    from eradiate_b200 import scenes
    from eradiate_b200.dist import ShardedRenderer
    from eradiate_b200.kernel import mi_load_dict, mi_traverse

    global SPP, WORKLOAD
    if args.config == "c3":
        SPP, WORKLOAD = C3_SPP, C3_WORKLOAD
        scene = mi_load_dict(scenes.config_c3(spp=SPP, res=C3_RES))
        npix = C3_RES * C3_RES
    else:
        scene = mi_load_dict(scenes.config_c2(spp=SPP, n_vza=N_VZA))
        npix = N_VZA
    wrapper = mi_traverse(scene)
    R = ShardedRenderer(scene, local_rank)
    paths_per_step_rank = npix * SPP
    seed = 20261017
    dev = torch.device(f"cuda:{local_rank}")
    flush_buf = torch.empty(512 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)  # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step(i):
        """HBM-resident step: this rank's shard + the all-reduce of the accumulators."""
        acc = R.launch(0, seed + i, SPP, sample_offset=rank * SPP)
        if world > 1:
            dist.all_reduce(acc, op=dist.ReduceOp.SUM)
        return acc

    # ---- K-bar (loop trips per path) from a stats-enabled run of the same workload ----
    stats = torch.zeros(8, dtype=torch.int64, device=dev)
    R.launch(0, seed, SPP, rank * SPP, stats=stats)
    torch.cuda.synchronize()
    sh = stats.cpu().numpy()
    kbar = float(sh[1] + sh[2]) / float(sh[0])
    bands = R.dev.render(0, seed, 16)[3].n_bands  # 1 = the reference's single global majorant

    for i in range(args.warmup):
        step(i)
    barrier()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    t_wall0 = time.perf_counter()
    for i in range(args.steps):
        flush_buf.fill_(float(i))  # L2 flush between timed iterations (not timed)
        ev[i][0].record()
        acc = R.accum(0)
        acc.zero_()
        kev[i][0].record()
        R.dev.render_device(0, seed + 1000 + i, SPP, rank * SPP, acc.data_ptr(), None,
                            torch.cuda.current_stream().cuda_stream)
        kev[i][1].record()
        if world > 1:
            dist.all_reduce(acc, op=dist.ReduceOp.SUM)
        ev[i][1].record()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    step_ms = sum(a.elapsed_time(b) for a, b in ev)
    kern_ms = sum(a.elapsed_time(b) for a, b in kev) / args.steps
    t = torch.tensor([step_ms, kern_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    step_ms, kern_ms = float(t[0]), float(t[1])
    ms_per_step = step_ms / args.steps
    value = world * paths_per_step_rank / (ms_per_step * 1e-3) / 1e6

    # ---- e2e: public host API, host buffers, H2D + D2H inside the timed region ----
    # the per-step host inputs: the sigma_t / albedo profiles the scene was built with
    sig = scene.flat._profile(scene.flat.medium.children["sigma_t"], 1200, "sigma_t")
    alb = scene.flat._profile(scene.flat.medium.children["albedo"], 1200, "albedo")
    sig_pinned = torch.from_numpy(sig.copy()).pin_memory()
    alb_pinned = torch.from_numpy(alb.copy()).pin_memory()
    upd = {
        "shape_atmosphere.interior_medium.sigma_t.volume.data": sig_pinned.numpy(),
        "shape_atmosphere.interior_medium.albedo.volume.data": alb_pinned.numpy(),
    }

    def e2e_step(i):
        wrapper.parameters.update(upd)          # host -> device (tables re-derived + uploaded)
        return R.render(0, seed + 5000 + i, world * SPP)  # render, all-reduce, device -> host

    for i in range(max(1, args.warmup // 2)):
        e2e_step(i)
    barrier()
    t0 = time.perf_counter()
    e2e_steps = args.steps
    for i in range(e2e_steps):
        out = e2e_step(i)
    barrier()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * paths_per_step_rank * e2e_steps / float(te[0]) / 1e6
    # table blob re-uploaded per step: sigma_t/majorant + albedo (float32 per layer)
    h2d_blob = int(R.dev.desc.n_layers * 4 * 2)
    d2h = 3 * npix * 8

    clocks = sampler.stop() if rank == 0 else None

    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        bytes_per_path = 2 * RECORD_BYTES * kbar + 16
        algo_bytes = paths_per_step_rank * bytes_per_path
        achieved = algo_bytes / (kern_ms * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {
                "workload": WORKLOAD, "paths_per_step_per_gpu": paths_per_step_rank,
                "parallelism": f"sample-sharded x{world}, one NCCL all-reduce of 3x{npix} f64 accumulators per step",
                "majorant_bands": int(bands),
                "l2": "flushed between timed iterations (512 MiB fill, outside the timed events)",
                "timing": "CUDA events on the launch stream per step, summed; max over ranks",
                "wall_s_timed_region": t_wall,
                "loop_trips_per_path": kbar,
            },
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_blob,
                    "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                    "api": "SceneParameters.update(host arrays) + ShardedRenderer.render() -> host float64"},
            "gpu_launches": args.steps,  # one ertb_render_kernel launch per step (per rank)
            "roofline": {
                "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": ncu_traffic_bytes(),
                "peak_source": peak_src,
                "algorithmic_bytes_per_path": bytes_per_path,
                "kernel_ms": kern_ms,
                "note": "algorithmic bytes = paths x (2 x 64 B x loop trips + 16 B) (SURVEY 8d); the megakernel "
                        "keeps the path record in registers, so real DRAM traffic is ~0 and the true bound is "
                        "the instruction-issue rate (see DESIGN.md, profiles/)",
            },
        }
        if args.config == "c2" and clocks and clocks.get("sm_mhz"):
            line["roofline"]["issue"] = ncu_issue(kern_ms, clocks["sm_mhz"],
                                                  torch.cuda.get_device_properties(0).multi_processor_count)
        if world == 1 and args.config == "c2":
            cores = os.cpu_count() or 1
            cpu_spp = 1 << 19  # 16.8 M paths: ~15-30 core-seconds of the same workload
            cpu_val, cpu_s, cpu_k = time_cpu_oracle(cpu_spp, repeats=2, n_threads=cores)
            line["cpu_baseline"] = {
                "value": cpu_val, "unit": UNIT, "cores": cores, "kind": "port",
                "sample": f"C2 scene, 32 pixels x spp=2^19 = {N_VZA * cpu_spp} paths, best of 2, "
                          f"{cpu_s:.2f} s per call, OpenMP on all {cores} host cores, oracle K={cpu_k:.2f}",
            }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--config", default="c2", choices=["c2", "c3"],
                    help="c2 = the headline (BASELINE configs[1]); c3 = configs[2], for the record only")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_cuda(args)


if __name__ == "__main__":
    main()
