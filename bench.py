#!/usr/bin/env python
"""
bench.py -- headline benchmark: Mpaths/s on the TOA-BRF AFGL1986+RPV scene
(BASELINE.json configs[1] "C2": AFGL-1986-shaped molecular atmosphere, RPV surface,
mono 550 nm, spherical shell, mdistant 32 VZA, spp = 2^20 per pixel per GPU).

    python bench.py --gpus N --steps K --warmup W          # CUDA path (this repo)
    python bench.py --impl reference --gpus N ...          # CPU arm: THE REFERENCE (oracle/_ref: Eradiate's
                                                           # Mitsuba fork, llvm_mono if a libLLVM is loadable,
                                                           # else scalar_mono, all host cores); the C oracle
                                                           # port only if oracle/_ref is absent

A "step" is one pass of the hot path: one render of the whole film (32 x 2^20 = 33.5 M
paths per GPU).  `value` is whole-job throughput with the scene tables resident in HBM
(device accumulators, CUDA-event timing on the launch stream, max over ranks, one NCCL
all-reduce of the accumulators inside the timed step when N > 1).  `e2e` is the same
metric through the boundary function Eradiate calls, `mi_render(mi_scene, [ctx], spp) ->
{spectral index: {sensor: Bitmap}}`: update map rendered on the host, tables pushed from host
memory, render, film developed on the host.  One JSON line is printed by rank 0.

Besides the headline the line carries (config.*): the same step sustained for >= 2 s with the clock
trace, the other BASELINE configurations (C1, C3 at full size, C4, one C5 band; N = 1 only), and
STRONG scaling at this N (total samples fixed: C2 at spp 2^20, C3 at spp 2^22, sample-sharded).
"""

from __future__ import annotations

import argparse
import hashlib
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
C3_RES = 32
C3_SPP_FULL = 1 << 22
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


def csrc_sha16() -> str:
    """Hash of the CUDA sources + compile flags: ties a committed ncu capture to the build it profiled."""
    h = hashlib.sha256()
    d = os.path.join(ROOT, "eradiate_b200", "csrc")
    for f in sorted(os.listdir(d)):
        if f.endswith((".cu", ".cuh")):
            h.update(open(os.path.join(d, f), "rb").read())
    h.update(open(os.path.join(ROOT, "include", "eradiate_b200.h"), "rb").read())
    return h.hexdigest()[:16]


def kernel_profile():
    """ncu figures of one C2 launch of the pool kernel (profiles/render_kernel_profile.json, written by
    tools/summarize_ncu.py --bench-json from the capture named inside), or None."""
    for name in ("render_kernel_profile.json", "render_kernel_traffic.json"):
        try:
            return json.load(open(os.path.join(ROOT, "profiles", name)))
        except Exception:
            continue
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
#        CPU arm: the reference itself (oracle/_ref), else the C oracle port
# ------------------------------------------------------------------------------


def time_cpu_oracle(spp: int, repeats: int = 1, n_threads: int = 0):
    """Returns (Mpaths/s, seconds per call, K-bar) of the C oracle PORT on a bounded C2 sample."""
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


class ReferenceArm:
    """The reference's own CPU implementation of the path: `mitsuba.render` of the identical C2
    dictionary through oracle/_ref -- `llvm_mono` when a libLLVM is loadable on this box (north star),
    else `scalar_mono`, the variant Eradiate's "mono" mode runs (_mode.py:56-123) -- on all host cores."""

    def __init__(self):
        from eradiate_b200 import scenes
        from oracle import ref

        self.cores = os.cpu_count() or 1
        self.ref = ref
        self.variant, self.why = "scalar_mono", "no loadable libLLVM on this box (ldconfig / DRJIT_LIBLLVM_PATH probed)"
        llvm = ref.llvm_runtime()
        if llvm and os.environ.get("ERTB_REF_VARIANT", "") != "scalar_mono":
            os.environ["DRJIT_LIBLLVM_PATH"] = llvm
            self.variant, self.why = "llvm_mono", f"libLLVM found: {llvm}"
        try:
            self.mi = ref.mitsuba(self.variant)
        except Exception as e:  # an LLVM the JIT cannot use after all
            self.variant, self.why = "scalar_mono", f"llvm_mono unusable here ({type(e).__name__}: {e})"[:200]
            self.mi = ref.mitsuba(self.variant)
        try:  # torchrun exports OMP_NUM_THREADS=1; the reference's own pool (nanothread) is sized explicitly
            import drjit

            drjit.set_thread_count(self.cores)
        except Exception:
            pass
        self.scene = self.mi.load_dict(ref.to_mitsuba(self.mi, scenes.config_c2(spp=64, n_vza=N_VZA)))

    def render(self, spp: int, seed: int) -> float:
        t0 = time.perf_counter()
        img = self.mi.render(self.scene, sensor=0, seed=seed, spp=spp)
        if self.variant.startswith("llvm"):
            import drjit

            drjit.eval(img)
            drjit.sync_thread()
        return time.perf_counter() - t0

    def describe(self) -> str:
        return (f"reference kernel built from /root/reference into oracle/_ref ({self.ref.describe()}), variant "
                f"{self.variant} ({self.why}), mitsuba.render of the C2 dict, {self.cores} host threads")


def reference_available() -> bool:
    try:
        from oracle import ref

        return ref.available()
    except Exception:
        return False


def run_reference(args):
    rank = env_int("RANK", 0)
    if rank != 0:
        return  # other ranks exit 0 without work
    cores = os.cpu_count() or 1
    if reference_available():
        arm = ReferenceArm()
        spp = 1 << 15  # bounded sample: 32 x 2^15 = 1.05 M paths per step
        for _ in range(args.warmup):
            arm.render(1 << 10, 1)
        times = [arm.render(spp, 100 + i) for i in range(args.steps)]
        kind, dtype, how = "reference", "f32", arm.describe()
    else:
        spp = 1 << 18  # the port is ~10x faster: 32 x 2^18 = 8.4 M paths per step
        for _ in range(args.warmup):
            time_cpu_oracle(1 << 12, n_threads=cores)  # n_threads explicit: torchrun exports OMP_NUM_THREADS=1
        times = [time_cpu_oracle(spp, n_threads=cores)[1] for _ in range(args.steps)]
        kind, dtype = "port", "f64"
        how = "oracle/_ref absent: C oracle port of the same algorithm (oracle/ertb_oracle.c), OpenMP"
    ms = 1e3 * float(np.mean(times))
    value = N_VZA * spp / (ms * 1e-3) / 1e6
    sample = (f"{args.steps} steps x (32 pixels x spp=2^{int(np.log2(spp))} = {N_VZA * spp} paths) of the C2 workload, "
              f"all {cores} host cores")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": dtype, "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": sample, "implementation": how},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample,
                         "implementation": how},
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
    from eradiate_b200.dist import ShardedRenderer, mi_render_sharded, shard_range
    from eradiate_b200.kernel import KernelContext, SeedState, mi_load_dict, mi_render, mi_traverse, render

    scene = mi_load_dict(scenes.config_c2(spp=SPP, n_vza=N_VZA))
    npix = N_VZA
    R = ShardedRenderer(scene, local_rank)
    paths_per_step_rank = npix * SPP
    seed = 20261017
    dev = torch.device(f"cuda:{local_rank}")
    flush_buf = torch.empty(512 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)  # > 126 MB L2
    n_sm = torch.cuda.get_device_properties(local_rank).multi_processor_count

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(*vals):
        t = torch.tensor(list(vals), dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(x) for x in t]

    def timed_steps(renderer, n_steps, spp_rank, offset, seed0, flush=True):
        """n_steps launches of this rank's shard (+ the all-reduce), CUDA events per step on the launch stream.
        Returns (sum of step ms, mean kernel-only ms), each the max over ranks."""
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n_steps)]
        kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n_steps)]
        barrier()
        for i in range(n_steps):
            if flush:
                flush_buf.fill_(float(i))  # L2 flush between timed iterations (outside the timed events)
            ev[i][0].record()
            acc = renderer.accum(0)
            acc.zero_()
            kev[i][0].record()
            renderer.dev.render_device(0, seed0 + i, spp_rank, offset, acc.data_ptr(), None,
                                       torch.cuda.current_stream().cuda_stream)
            kev[i][1].record()
            if world > 1:
                dist.all_reduce(acc, op=dist.ReduceOp.SUM)
            ev[i][1].record()
        barrier()
        step_ms = sum(a.elapsed_time(b) for a, b in ev)
        kern_ms = sum(a.elapsed_time(b) for a, b in kev) / n_steps
        return max_over_ranks(step_ms, kern_ms)

    # ---- K-bar (loop trips per path) from a stats-enabled run of the same workload ----
    stats = torch.zeros(8, dtype=torch.int64, device=dev)
    R.launch(0, seed, SPP, rank * SPP, stats=stats)
    torch.cuda.synchronize()
    sh = stats.cpu().numpy()
    kbar = float(sh[1] + sh[2]) / float(sh[0])
    bands = R.dev.render(0, seed, 16)[3].n_bands  # 1 = the reference's single global majorant

    for i in range(args.warmup):
        R.launch(0, seed + i, SPP, sample_offset=rank * SPP)
    barrier()

    # ---- headline: K timed steps ---------------------------------------------------------------------
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    t_wall0 = time.perf_counter()
    step_ms, kern_ms = timed_steps(R, args.steps, SPP, rank * SPP, seed + 1000)
    t_wall = time.perf_counter() - t_wall0
    ms_per_step = step_ms / args.steps
    value = world * paths_per_step_rank / (ms_per_step * 1e-3) / 1e6

    # ---- e2e: the boundary function, host buffers in, developed Bitmap out, every step -----------------
    # mi_render(mi_scene, [ctx], spp): renders the update map for the context on the host (sigma_t / albedo
    # profiles + irradiance -> host arrays), pushes them (H2D), renders, reads the accumulators back (D2H) and
    # develops the film on the host.  N > 1: the same call sample-sharded over the ranks (one all-reduce).
    mi_scene = mi_traverse(scene, scenes.spectral_update_map(R.dev.desc.n_layers, spherical=True))
    ctx = KernelContext(w=550.0)

    def e2e_step(i):
        if world > 1:
            res = mi_render_sharded(mi_scene, [ctx], spp=world * SPP, seed_state=SeedState(i), shard="samples")
        else:
            res = mi_render(mi_scene, [ctx], spp=SPP, seed_state=SeedState(i))
        bmp = res[ctx.si.as_hashable]["measure"]
        return np.array(bmp)  # the developed film, host memory

    for i in range(max(1, args.warmup // 2)):
        e2e_step(i)
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        film = e2e_step(100 + i)
    barrier()
    (e2e_s,) = max_over_ranks(time.perf_counter() - t0)
    e2e_value = world * paths_per_step_rank * args.steps / e2e_s / 1e6
    h2d = int(R.dev.desc.n_layers * 4 * 2 + 4)  # sigma_t + albedo (float32 per layer) + irradiance
    d2h = 3 * npix * 8
    clocks = sampler.stop() if rank == 0 else None

    # ---- sustained: the same step back to back for >= 2 s (no L2 flush: steps are 3.7 ms of compute) ----
    sus = ClockSampler(local_rank)
    if rank == 0:
        sus.start()
    n_sus = max(50, int(2200.0 / max(kern_ms, 0.05)))
    sus_ms, sus_kern = timed_steps(R, n_sus, SPP, rank * SPP, seed + 9000, flush=False)
    sus_clocks = sus.stop() if rank == 0 else None
    sustained = {"steps": n_sus, "seconds": sus_ms * 1e-3, "ms_per_step": sus_ms / n_sus,
                 "value": world * paths_per_step_rank / (sus_ms / n_sus * 1e-3) / 1e6, "unit": UNIT,
                 "clocks": sus_clocks}

    # ---- strong scaling at this N: total samples fixed, sample-sharded -------------------
    strong = {}
    off, cnt = shard_range(SPP, rank, world)
    s_ms, s_kern = timed_steps(R, 10, cnt, off, seed + 20000)
    strong["c2"] = {"workload": "C2, spp 2^20 per pixel IN TOTAL, split over the ranks", "ms_per_step": s_ms / 10,
                    "kernel_ms": s_kern, "value": npix * SPP / (s_ms / 10 * 1e-3) / 1e6, "unit": UNIT}
    sc3 = mi_load_dict(scenes.config_c3(spp=C3_SPP_FULL, res=C3_RES))
    R3 = ShardedRenderer(sc3, local_rank)
    off, cnt = shard_range(C3_SPP_FULL, rank, world)
    R3.launch(0, seed, max(cnt >> 6, 16), off)
    s_ms, s_kern = timed_steps(R3, 3, cnt, off, seed + 30000, flush=False)
    strong["c3"] = {"workload": "C3 (BASELINE configs[2]): AFGL-shaped + aerosol layer (tabphase), hdistant 32x32, "
                                "spp 2^22 per pixel IN TOTAL (4.3e9 paths), split over the ranks",
                    "ms_per_step": s_ms / 3, "kernel_ms": s_kern,
                    "value": C3_RES * C3_RES * C3_SPP_FULL / (s_ms / 3 * 1e-3) / 1e6, "unit": UNIT,
                    "majorant_bands": int(R3.dev.render(0, seed, 16)[3].n_bands)}
    del R3

    # ---- the other BASELINE configurations (rank 0 of a 1-GPU run only) --------------------------------
    others = None
    if world == 1:
        def run_cfg(label, kd, spp_, sensor=0, steps=3):
            sc_ = mi_load_dict(kd)
            render(sc_, sensor=sensor, seed=1, spp=max(16, spp_ >> 6))  # tables / BVH upload, warm-up
            ms, st = [], None
            for r_ in range(steps):
                st = render(sc_, sensor=sensor, seed=2 + r_, spp=spp_).stats
                ms.append(st["device_ms"])
            return {"config": label, "paths_per_step": st["n_paths"], "steps": steps, "ms_per_step": float(np.mean(ms)),
                    "value": st["n_paths"] / float(np.mean(ms)) / 1e3, "unit": UNIT,
                    "loop_trips_per_path": (st["trips_main"] + st["trips_nee"]) / st["n_paths"],
                    "majorant_bands": st["n_bands"]}

        c4 = scenes.config_c4(spp=16)
        others = [
            run_cfg("C1: homogeneous + Lambertian, plane-parallel, mdistant 1 angle, spp 4096 x 4096 launches' worth",
                    scenes.config_c1(), 4096 * 4096),
            run_cfg("C3 full size: hdistant 32x32, spp 2^22 (4.3e9 paths)", scenes.config_c3(), C3_SPP_FULL),
            run_cfg("C4: disc canopy (LAI 3, 60k leaves x 25 instances) + AFGL + RPV, mdistant 32, spp 2^18", c4, 1 << 18),
            run_cfg("C4: same scene, perspective 64x64 inside the atmosphere, spp 2^11", c4, 1 << 11, sensor=1),
            run_cfg("C5 band @550 nm: polarized ocean + AFGL + polarized aerosol, spherical, mdistant 1, spp 2^26",
                    scenes.config_c5(), 1 << 26),
            # a canopy of small groups (3 abstract trees: 300 disc leaves + a trunk each): the VOTE form of the 3D
            # kernel's BVH stage, as opposed to C4's 60 k-leaf groups
            run_cfg("abstract-tree canopy (3 trees x 300 leaves + trunks) + AFGL + Lambertian, mdistant 5, spp 2^20",
                    scenes.atmosphere_scene(geometry="plane_parallel", n_layers=60, sza=40.0, saa=30.0,
                                            canopy={"trees": {}, "size": (8.0, 8.0, 4.1)},
                                            surface={"type": "diffuse", "reflectance": 0.2},
                                            sensor={"type": "mdistant", "vza": [-55.0, -20.0, 0.0, 30.0, 65.0], "vaa": 30.0}),
                    1 << 20),
        ]

    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        bytes_per_path = 2 * RECORD_BYTES * kbar + 16
        hbm_achieved = paths_per_step_rank * bytes_per_path / (kern_ms * 1e-3) / 1e9
        prof = kernel_profile() or {}
        sm_mhz = (clocks or {}).get("sm_mhz") or (clocks or {}).get("sm_max_mhz") or 1965.0
        issue_peak = n_sm * 4 * float(sm_mhz) * 1e6 / 1e9  # 4 schedulers per SM, 1 warp instruction per clock
        roofline = {"bound": "issue", "unit": "G warp-inst/s", "peak": issue_peak, "achieved": None, "frac": None,
                    "traffic": prof.get("dram_bytes_per_launch"), "kernel_ms": kern_ms,
                    "peak_source": f"{n_sm} SMs x 4 schedulers x {sm_mhz:.0f} MHz (clock sampled during the timed region)"}
        if prof.get("warp_inst_per_launch"):
            inst = float(prof["warp_inst_per_launch"])
            lanes = float(prof.get("thread_inst_per_warp_inst", 0.0))
            roofline.update({
                "achieved": inst / (kern_ms * 1e-3) / 1e9, "frac": inst / (kern_ms * 1e-3) / 1e9 / issue_peak,
                "active_lanes": lanes, "thread_issue_frac": inst / (kern_ms * 1e-3) / 1e9 / issue_peak * lanes / 32.0,
                "instructions": "smsp__inst_executed.sum of one launch of this workload (ncu capture "
                                f"{prof.get('capture', 'profiles/')}), divided by the kernel time measured live",
                "profile_matches_build": prof.get("csrc_sha16") == csrc_sha16(),
            })
        roofline["hbm_algorithmic"] = {
            "achieved": hbm_achieved, "peak": peak, "unit": "GB/s", "frac": hbm_achieved / peak, "peak_source": peak_src,
            "algorithmic_bytes_per_path": bytes_per_path,
            "note": "SURVEY 8d figure for an HBM wavefront design: paths x (2 x 64 B x loop trips + 16 B) / kernel time. "
                    "This kernel keeps path records in shared memory (traffic = measured DRAM bytes per launch), so "
                    "the figure exceeds the HBM peak: it is the speed relative to that design's ceiling, not a utilisation",
        }
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
                "sustained": sustained,
                "strong": strong,
                "other_configs": others,
            },
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "steps": args.steps,
                    "api": "mi_render(mi_scene, [ctx], spp) -> {si: {sensor: Bitmap}} (eradiate.kernel boundary); "
                           "N > 1: dist.mi_render_sharded(..., shard='samples')",
                    "film_checksum": float(np.asarray(film, dtype=np.float64).sum())},
            "gpu_launches": args.steps,  # one render kernel launch per timed step (per rank)
            "roofline": roofline,
        }
        if world == 1:
            cores = os.cpu_count() or 1
            if reference_available():
                arm = ReferenceArm()
                arm.render(1 << 10, 1)
                cpu_spp = 1 << 17  # 4.2 M paths of the same workload: ~10-30 s of all host cores
                t_cpu = min(arm.render(cpu_spp, 7), arm.render(cpu_spp, 8))
                line["cpu_baseline"] = {
                    "value": N_VZA * cpu_spp / t_cpu / 1e6, "unit": UNIT, "cores": cores, "kind": "reference",
                    "sample": f"C2 scene, 32 pixels x spp=2^17 = {N_VZA * cpu_spp} paths, best of 2, {t_cpu:.2f} s per call",
                    "implementation": arm.describe(),
                }
            else:
                cpu_spp = 1 << 19
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
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_cuda(args)


if __name__ == "__main__":
    main()
