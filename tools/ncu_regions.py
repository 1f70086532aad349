#!/usr/bin/env python
"""Break an .ncu-rep of the pool kernel down by code region (enclosing device function / phase of the kernel):
share of warp instructions, active lanes, and instruction mix (FP32 pipe, MUFU, integer/logic, shared memory).
Usage: python tools/ncu_regions.py gpurun_out/r01d_pool.ncu-rep profiles/r01d_pool_regions.md <dir with the csrc files of the
profiled commit>   (git show <commit>:eradiate_b200/csrc/<file> > dir/<file>)"""
import collections, csv, io, re, subprocess, sys

import os
rep, out, srcdir = sys.argv[1], sys.argv[2], sys.argv[3]
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
BLOCKS = {
    "BSDF block (surface_interact, rpv/rtls/hapke/ocean eval, cosine-hemisphere sampling)":
        r"surface_interact|rpv_eval|rtls_eval|hapke|bsdf_f|cos_dphi|cosine_hemisphere|disk_concentric|\bonb\b|oc_\w+|uniform_hemisphere",
    "phase functions (eval + sampling, tabulated CDF search)": r"leaf_eval|leaf_sample|tab_eval|tab_sample|rayleigh_|hg_eval",
    "RNG (PCG32)": r"pcg_next|pcg_float|pcg_seed|mix64",
    "altitude + layer lookup (walk)": r"altitude_at|layer_of",
    "segment set-up / primary rays": r"segment_setup|primary_entry|band_exit|band_of",
    "vector helpers (dot3, fma3, ...)": r"\bdot3\b|\bfma3\b|scale3|normalize3|safe_sqrtf|clampf|fast_sqrt|\bmk3\b",
}
FUNC = re.compile(r"__device__[^;(]*?\b([A-Za-z_][A-Za-z_0-9]*)\s*\(")
FP32 = re.compile(r"^(FFMA|FMUL|FADD|FMNMX|FSEL|FSET|FSETP|FCHK|F2F|I2F|F2I|FRND|HFMA2|HADD2|HMUL2)")
MUFU = re.compile(r"^MUFU")
INT = re.compile(r"^(IMAD|IADD|LOP|SHF|LEA|ISETP|SEL|MOV|PRMT|POPC|FLO|BREV|IABS|IMNMX|VOTE|SHFL|S2R|CS2R|R2P|P2R|PLOP|UMOV|ULOP|UIADD|USHF|UISETP|UIMAD|ULEA|USEL|REDUX|MATCH|I2I|BMSK|SGXT|LDC|ULDC|R2UR|S2UR)")
MEM = re.compile(r"^(LDS|STS|LDG|STG|LD\b|ST\b|ATOM|RED|LDL|STL|UBLKCP|SYNCS|MEMBAR|CCTL|ERRBAR|FENCE)")

def region_map(fname):
    """line number -> region, from the source file of the profiled commit"""
    path = os.path.join(srcdir, fname)
    if not os.path.exists(path):
        return {}
    out_, region, depth_fn = {}, f"other ({fname})", None
    for ln, text in enumerate(open(path).read().splitlines(), 1):
        if fname == "ertb_kernel_pool.cuh":
            if "__global__" in text: region = "pool kernel: prologue / epilogue"
            if "membership masks" in text and "//" in text and ln > 100: region = "pool kernel: warp scheduler (ballots, phase choice, compaction)"
            elif "free-flight walk" in text and ln > 100: region = "pool kernel: walk phase body (record loads/stores, loop control, tracking logic)"
            elif "finish the previous path" in text: region = "pool kernel: finish + regenerate phase"
            elif "heavy events" in text: region = "pool kernel: surface / scatter event phases"
            elif text.startswith("#undef FLD"): region = "pool kernel: prologue / epilogue"
            elif "film_flush" in text and "__device__" in text: region = "film flush"
        m = FUNC.search(text)
        if m and not (fname == "ertb_kernel_pool.cuh" and "__global__" in text):
            name = m.group(1)
            hit = next((k for k, pat in BLOCKS.items() if re.search(pat, name)), None)
            if hit: region = hit
            elif fname != "ertb_kernel_pool.cuh": region = f"other ({fname})"
        out_[ln] = region
    return out_


maps = {}
cur_file = None
stats = collections.defaultdict(lambda: collections.Counter())
region = None
for r in csv.reader(io.StringIO(src)):
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        if cur_file not in maps:
            maps[cur_file] = region_map(cur_file)
        continue
    if r[0].isdigit():
        region = maps[cur_file].get(int(r[0]), f"other ({cur_file})")
        continue
    if r[0] == "" and len(r) > 8 and r[3].strip():  # a SASS row under the current source line
        try:
            ie, tie = int(r[7]), int(r[8])
        except ValueError:
            continue
        op = re.sub(r"^@!?U?P\d+\s+", "", r[3].strip()).split()[0]
        s_ = stats[region]
        s_["inst"] += ie
        s_["thread"] += tie
        s_["fp32" if FP32.match(op) else "mufu" if MUFU.match(op) else "mem" if MEM.match(op) else "int" if INT.match(op) else "ctl"] += ie
tot = sum(s["inst"] for s in stats.values()) or 1
with open(out, "w") as f:
    f.write(f"# Code regions of `{rep.split('/')[-1]}`\n\nWarp instructions executed, by enclosing device function / phase of the kernel "
            "(ncu source page, SASS rows attributed to their CUDA source line). `lanes` = thread instructions / warp "
            "instructions. Mix columns are shares of the region's own warp instructions: FP32 pipe (FFMA/FMUL/FADD/"
            "FMNMX/conversions), MUFU (log/sqrt/rcp/sin/cos/ex2), integer + logic + predicates, memory (shared/global/atomics), "
            "control (branches, barriers, warp sync).\n\n")
    f.write("| region | share | lanes | FP32 | MUFU | int | mem | ctl |\n|---|---|---|---|---|---|---|---|\n")
    for k, s in sorted(stats.items(), key=lambda kv: -kv[1]["inst"]):
        n = s["inst"] or 1
        f.write(f"| {k} | {100 * s['inst'] / tot:.1f}% | {s['thread'] / n:.1f} | {100 * s['fp32'] / n:.0f}% | {100 * s['mufu'] / n:.0f}% | "
                f"{100 * s['int'] / n:.0f}% | {100 * s['mem'] / n:.0f}% | {100 * s['ctl'] / n:.0f}% |\n")
print(open(out).read())
