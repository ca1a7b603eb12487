"""Per-kernel SASS opcode histogram and ptxas resource summary of libhpdecode.so, for profiles/.

usage: python tools/sass_summary.py profiles/rN   ->   profiles/rN_sass_opcodes.txt, profiles/rN_ptxas_summary.txt
Needs only the CUDA toolkit (cuobjdump, c++filt): runs in the build container, no GPU.
"""
import collections
import glob
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "pytorch-human-pose_b200", "hpdecode", "libhpdecode.so")
CSRC = os.path.join(ROOT, "pytorch-human-pose_b200", "csrc")
# mnemonics worth a column of their own: global / shared memory width, TMA, tensor cores, FMA, shuffles, votes
KEY = ["LDG.E.128", "LDG.E.64", "LDG.E", "STG.E.128", "STG.E.64", "STG.E", "LDS", "STS", "FFMA", "FMUL", "FADD", "FMNMX", "DFMA",
       "SHFL", "VOTE", "REDUX", "BAR", "ATOMG", "UTMALDG", "UTMASTG", "UTCMMA", "HMMA"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    short = []
    for n in out:
        n = re.sub(r"\(anonymous namespace\)::", "", n)
        n = re.sub(r"^void ", "", n)
        n = re.sub(r"hpd::", "", n)
        n = re.sub(r"\(int\)", "", n)
        n = re.sub(r"\(.*\)$", "", n)
        short.append(n)
    return short


def sass_histogram(path):
    txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for line in txt.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = kernels.setdefault(m.group(1), collections.Counter())
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur is not None:
            cur[m.group(1)] += 1
    names = demangle(list(kernels))
    with open(path, "w") as f:
        f.write("cuobjdump -sass libhpdecode.so (sm_100a): static instruction counts per kernel; columns = selected opcode "
                "prefixes, 'all' = every instruction.  No UTMALDG/UTMASTG/UTCMMA/HMMA: the path stages through registers "
                "and shared memory and has no dense contraction (DESIGN.md 3).\n")
        f.write("%-48s %7s " % ("kernel", "all") + " ".join("%9s" % k for k in KEY) + "\n")
        for name, (_, c) in sorted(zip(names, kernels.items())):
            tot = sum(c.values())
            cols = []
            for k in KEY:
                n = sum(v for op, v in c.items() if op == k or op.startswith(k + ".")) if k not in ("LDG.E", "STG.E") else \
                    sum(v for op, v in c.items() if (op == k or op.startswith(k + ".")) and ".128" not in op and ".64" not in op)
                cols.append(n)
            f.write("%-48s %7d " % (name[:48], tot) + " ".join("%9d" % n for n in cols) + "\n")
    return len(kernels)


def ptxas_summary(path):
    rows = []
    for log in sorted(glob.glob(os.path.join(CSRC, "*.ptxas.log"))):
        lines = open(log).read().splitlines()
        for i, ln in enumerate(lines):
            m = re.search(r"Compiling entry function '(\S+)'", ln)
            if not m:
                continue
            props = " | ".join(x.strip().replace("ptxas info    : ", "") for x in lines[i + 2:i + 4])
            rows.append((m.group(1), props))
    names = demangle([r[0] for r in rows])
    with open(path, "w") as f:
        f.write("nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -Xptxas -v: stack / spills | registers, "
                "barriers, static smem per kernel\n")
        for n, (_, props) in sorted(zip(names, rows)):
            f.write("%-48s %s\n" % (n[:48], props))
    return len(rows)


if __name__ == "__main__":
    prefix = sys.argv[1]
    print(sass_histogram(prefix + "_sass_opcodes.txt"), "kernels in SASS;", ptxas_summary(prefix + "_ptxas_summary.txt"), "in ptxas logs")
