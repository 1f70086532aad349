"""Summarise an .ncu-rep (render kernel) into profiles/<name>.md + .json.  Usage:
   python tools/summarize_ncu.py gpurun_out/r01_prof.ncu-rep profiles/r01_render_kernel [note]"""
import csv, io, json, subprocess, sys, collections

rep, out = sys.argv[1], sys.argv[2]
note = sys.argv[3] if len(sys.argv) > 3 else ""
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
KEYS = [
    "Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_xu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_xu.sum",
    "sm__inst_executed_pipe_lsu.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_elapsed.max",
]
kernels = []
for r in rows[2:]:
    k = {}
    for key in KEYS:
        if key in hdr:
            i = hdr.index(key)
            k[key] = (r[i], units[i])
    for i, h in enumerate(hdr):
        if "warp_issue_stalled" in h and h.endswith("per_warp_active.pct"):
            try:
                if float(r[i]) > 2.0:
                    k[h] = (r[i], "%")
            except ValueError:
                pass
    kernels.append(k)

# hottest source lines
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
d = collections.OrderedDict()
cur = None
for r in csv.reader(io.StringIO(src)):
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if r[0].isdigit() and len(r) > 8:
        try:
            key = (cur, int(r[0]))
            v = d.setdefault(key, [r[1].strip()[:110], 0, 0])
            v[1] += int(r[7]); v[2] += int(r[8])
        except ValueError:
            pass
tot = sum(v[1] for v in d.values()) or 1
top = sorted(d.items(), key=lambda kv: -kv[1][1])[:25]

with open(out + ".md", "w") as f:
    f.write(f"# ncu summary: `{rep.split('/')[-1]}`\n\n{note}\n\n")
    f.write("Captured with `ncu --set full --clock-control none --import-source on -k regex:<kernel>` "
            "under gpurun (1x B200); numbers under the profiler are cold-cache/serialised -- use shares, not absolutes.\n\n")
    for n, k in enumerate(kernels):
        f.write(f"## launch {n}\n\n| metric | value | unit |\n|---|---|---|\n")
        for key, (v, u) in k.items():
            f.write(f"| `{key}` | {v} | {u} |\n")
        f.write("\n")
    f.write("## hottest CUDA source lines (share of warp instructions executed, avg active threads)\n\n")
    f.write("| file:line | share | thr/inst | source |\n|---|---|---|---|\n")
    for (fl, ln), (s, ie, tie) in top:
        f.write(f"| {fl}:{ln} | {100*ie/tot:.1f}% | {tie/max(ie,1):.1f} | `{s.replace('|', '/')}` |\n")
json.dump({"report": rep, "kernels": [{k: v[0] for k, v in kk.items()} for kk in kernels]}, open(out + ".json", "w"), indent=1)
if "--bench-json" in sys.argv:
    # profiles/render_kernel_profile.json: what bench.py's issue roofline reads (instruction count of one launch of the
    # headline workload), tied to the build it was captured from by the hash of the CUDA sources
    import hashlib, os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    h = hashlib.sha256()
    d = os.path.join(root, "eradiate_b200", "csrc")
    for f in sorted(os.listdir(d)):
        if f.endswith((".cu", ".cuh")):
            h.update(open(os.path.join(d, f), "rb").read())
    h.update(open(os.path.join(root, "include", "eradiate_b200.h"), "rb").read())
    k0 = kernels[0]
    json.dump({
        "capture": out + ".md", "kernel": k0["Kernel Name"][0], "csrc_sha16": h.hexdigest()[:16],
        "warp_inst_per_launch": float(k0["smsp__inst_executed.sum"][0]),
        "thread_inst_per_warp_inst": float(k0["smsp__thread_inst_executed_per_inst_executed.ratio"][0]),
        "dram_bytes_per_launch": float(k0["dram__bytes_read.sum"][0]) * {"Kbyte": 1e3, "Mbyte": 1e6, "byte": 1.0, "Gbyte": 1e9}[k0["dram__bytes_read.sum"][1]]
                                 + float(k0["dram__bytes_write.sum"][0]) * {"Kbyte": 1e3, "Mbyte": 1e6, "byte": 1.0, "Gbyte": 1e9}[k0["dram__bytes_write.sum"][1]],
        "issue_active_pct": float(k0["smsp__issue_active.avg.pct_of_peak_sustained_active"][0]),
        "ncu_duration_ms": float(k0["gpu__time_duration.sum"][0]),
    }, open(os.path.join(root, "profiles", "render_kernel_profile.json"), "w"), indent=1)
    print("wrote profiles/render_kernel_profile.json")
print("wrote", out + ".md")
