"""Opcode census of the built library: `cuobjdump -sass reart_b200/csrc/libreart_b200.so`, per kernel, for the
instructions that prove what the binary is made of (VERDICT r01 "No SASS listing is committed").

    python scripts/sass_summary.py > profiles/r02_sass_summary.md
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "reart_b200", "csrc", "libreart_b200.so")
COLS = ["UBLKCP", "SYNCS", "FFMA2", "FADD2", "FMUL2", "FFMA", "FMNMX3", "FMNMX", "REDUX", "SHFL", "BAR", "RED", "ATOM", "DADD",
        "MUFU", "LDG", "STG", "LDS", "STS", "STL", "LDL", "UTCMMA|HGMMA|HMMA", "float RED/ATOM"]


def census(so=SO):
    out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
    arch = sorted(set(re.findall(r"arch = (sm_\w+)", out)))
    kernels = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)([A-Z0-9_.]*)", line)
        if m and cur:
            op, suffix = m.group(1), m.group(2)
            kernels[cur]["_total"] += 1
            for c in COLS[:-1]:
                pat = "C?REDUX" if c == "REDUX" else c
                if re.fullmatch("(" + pat + r")\w*", op):
                    kernels[cur][c] += 1
            if re.fullmatch(r"(RED|ATOM)\w*", op) and re.search(r"\.F(16|32|64)", suffix):
                kernels[cur]["float RED/ATOM"] += 1
    return arch, kernels


def demangle(names):
    try:
        out = subprocess.run(["cu++filt"] + names, capture_output=True, text=True, check=True).stdout.splitlines()
        return [re.sub(r"\(.*$", "", o.replace("(int)", "").replace("(bool)", "")).replace("void ", "").replace("reart::", "") for o in out]
    except Exception:
        return names


def main():
    arch, kernels = census()
    names = list(kernels)
    short = demangle(names)
    print("# r02 -- SASS opcode census of `reart_b200/csrc/libreart_b200.so` (`cuobjdump -sass`, `scripts/sass_summary.py`)\n")
    print(f"Target architectures in the fatbin: **{', '.join(arch)}**.  Counts are static instruction counts per kernel "
          "(opcode prefix match; e.g. `RED` = `REDG`/`RED.E...`, `REDUX` = `CREDUX`/`REDUX`).  No tensor-core opcode "
          "(`UTCMMA`/`HGMMA`/`HMMA`) appears anywhere: the contraction depth of the path is 3 (DESIGN.md 4.1).\n")
    print("| kernel | instr | " + " | ".join(COLS) + " |")
    print("|---|---:|" + "---:|" * len(COLS))
    for n, s in sorted(zip(names, short), key=lambda kv: -kernels[kv[0]]["_total"]):
        k = kernels[n]
        print(f"| `{s[:60]}` | {k['_total']} | " + " | ".join(str(k[c]) if k[c] else "" for c in COLS) + " |")
    tot = collections.Counter()
    for k in kernels.values():
        tot.update(k)
    print(f"| **all kernels** | {tot['_total']} | " + " | ".join(str(tot[c]) if tot[c] else "" for c in COLS) + " |")


if __name__ == "__main__":
    main()
