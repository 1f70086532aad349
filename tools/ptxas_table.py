"""Condenses an `nvcc -Xptxas=-v` log into one line per kernel: registers, spills, stack, shared memory.
Usage: python tools/ptxas_table.py LOG [LOG_BEFORE]  (with two logs: only the kernels that changed)"""
import re
import subprocess
import sys


def parse(path):
    out, name = {}, None
    for line in open(path, errors="replace"):
        m = re.search(r"Compiling entry function '(\S+)'", line)
        if m:
            name = m.group(1)
            out[name] = {}
            continue
        if name is None:
            continue
        m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
        if m:
            out[name].update(stack=int(m.group(1)), spill_st=int(m.group(2)), spill_ld=int(m.group(3)))
        m = re.search(r"Used (\d+) registers", line)
        if m:
            out[name]["regs"] = int(m.group(1))
            s = re.search(r"(\d+) bytes smem", line)
            out[name]["smem"] = int(s.group(1)) if s else 0
    return out


def demangle(names):
    r = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True)
    return dict(zip(names, r.stdout.splitlines()))


def main():
    new = parse(sys.argv[1])
    old = parse(sys.argv[2]) if len(sys.argv) > 2 else None
    dm = demangle(list(new))
    for k, v in new.items():
        if old is not None and old.get(k) == v:
            continue
        short = re.sub(r"\(.*", "", dm[k])
        was = f"   (was {old[k]})" if old is not None and k in old else ""
        print(f"{short:70s} {v}{was}")


if __name__ == "__main__":
    main()
