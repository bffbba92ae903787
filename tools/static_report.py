#!/usr/bin/env python
"""Static evidence from the build, no GPU needed: registers / stack / spills per kernel (ptxas -v logs written by the
Makefile) and the tensor-core / TMA / async-copy SASS mnemonics per kernel of the shipped libtnml.so (cuobjdump -sass).

    python tools/static_report.py > profiles/r02_static_ptxas_sass.txt
"""
import glob
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "tensornetworkforml_b200", "csrc", "build")
LIB = os.path.join(ROOT, "tensornetworkforml_b200", "lib", "libtnml.so")
# what each mnemonic proves (B200_PROFILING.md): DMMA = FP64 mma.sync, UTCHMMA = tcgen05.mma (kind::tf32/f16),
# LDTM = tcgen05.ld (TMEM -> registers), UTMALDG = cp.async.bulk.tensor (TMA tile load), UBLKCP = cp.async.bulk,
# LDGSTS = cp.async, SYNCS = mbarrier, UCGABAR = cluster barrier, REDUX = warp reduction in the integer pipe
MNEMONICS = ["DMMA", "UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UBLKCP", "LDGSTS", "SYNCS", "UCGABAR", "REDUX",
             "DFMA", "FFMA", "MUFU"]


def demangle(names):
    out = subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.splitlines()
    return [re.sub(r"\(.*", "", d) for d in out]


def ptxas_rows():
    rows = []
    for f in sorted(glob.glob(os.path.join(BUILD, "*.ptxas.log"))):
        name, frame = None, ("0", "0", "0")
        for ln in open(f):
            m = re.search(r"Compiling entry function '(\S+)' for 'sm_100a'", ln)
            if m:
                name, frame = m.group(1), ("0", "0", "0")
                continue
            m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", ln)
            if name and m:
                frame = m.groups()
            m = re.search(r"Used (\d+) registers", ln)
            if name and m:
                sm = re.search(r"(\d+) bytes smem", ln)
                rows.append((os.path.basename(f).split(".")[0], name, int(m.group(1)), frame, sm.group(1) if sm else "0"))
                name = None
    return rows


def sass_counts():
    txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    counts, cur = {}, None
    for ln in txt.splitlines():
        m = re.match(r"\s*Function : (\S+)", ln)
        if m:
            cur = counts.setdefault(m.group(1), dict.fromkeys(MNEMONICS, 0))
            cur["_total"] = 0
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", ln)
        if m:
            cur["_total"] += 1
            op = m.group(1)
            if op in cur:
                cur[op] += 1
    return counts


def main():
    rows = ptxas_rows()
    names = demangle([r[1] for r in rows])
    print("# registers / stack frame / spills per kernel (nvcc 12.9, -O3, sm_100a; static shared memory only)")
    print("%-12s %-44s %5s %6s %9s %9s %7s" % ("file", "kernel", "regs", "stack", "spill st", "spill ld", "smem"))
    for (f, _, r, fr, sm), d in zip(rows, names):
        print("%-12s %-44s %5d %6s %9s %9s %7s" % (f, d.replace("tnml::", "").replace("void ", "")[:44], r, fr[0], fr[1],
                                                 fr[2], sm))
    if not os.path.exists(LIB):
        print("\n(libtnml.so not built: no SASS table)")
        return
    counts = sass_counts()
    dn = demangle(list(counts))
    cols = [m for m in MNEMONICS if any(c[m] for c in counts.values())]
    print("\n# SASS mnemonics per kernel of tensornetworkforml_b200/lib/libtnml.so (cuobjdump -sass); instr = all instructions")
    print("%-44s %7s " % ("kernel", "instr") + " ".join("%7s" % c for c in cols))
    for (k, c), d in sorted(zip(counts.items(), dn), key=lambda t: t[1]):
        if not any(c[m] for m in cols if m not in ("DFMA", "FFMA", "MUFU")):
            continue                               # only kernels that use tensor cores, TMA, async copies or mbarriers
        print("%-44s %7d " % (d.replace("tnml::", "").replace("void ", "")[:44], c["_total"]) +
              " ".join("%7d" % c[m] for m in cols))
    tot = {m: sum(c[m] for c in counts.values()) for m in cols}
    print("%-44s %7d " % ("all %d kernels" % len(counts), sum(c["_total"] for c in counts.values())) +
          " ".join("%7d" % tot[m] for m in cols))


if __name__ == "__main__":
    sys.exit(main())
