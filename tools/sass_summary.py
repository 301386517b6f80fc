#!/usr/bin/env python
"""Per-kernel SASS opcode summary of libdvid_b200.so: which kernels run on the Blackwell tensor path.

  python tools/sass_summary.py [sass listing file] > profiles/r02_sass_opcodes.txt     (runs cuobjdump -sass itself)

Mnemonics (guides/B200_PROFILING.md): UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG = TMA tensor
load/store, UBLKCP/UBLKPF = cp.async.bulk / L2 prefetch, SYNCS = mbarrier, HMMA = mma.sync (legacy warp-level tensor
path), LDGSTS = cp.async."""
import re
import subprocess
import sys

KEYS = ["UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UBLKPF", "SYNCS", "HMMA", "LDGSTS", "MUFU"]


def main():
    if len(sys.argv) > 1:
        src = open(sys.argv[1]).read()
    else:
        src = subprocess.run(["cuobjdump", "-sass", "diffusionvid_b200/_C/libdvid_b200.so"], capture_output=True,
                             text=True, stdin=subprocess.DEVNULL).stdout
    out = {}
    name = None
    for ln in src.splitlines():
        m = re.match(r"\s*Function : (\S+)", ln)
        if m:
            name = m.group(1)
            out[name] = dict.fromkeys(KEYS, 0)
            out[name]["_n"] = 0
            continue
        if name is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
        if not m:
            continue
        op = m.group(1)
        out[name]["_n"] += 1
        for k in KEYS:
            if op.startswith(k):
                out[name][k] += 1
    demangled = subprocess.run(["c++filt"], input="\n".join(out), capture_output=True, text=True).stdout.splitlines()
    print("%-72s %6s " % ("kernel (sm_100a SASS)", "instr") + " ".join("%7s" % k for k in KEYS))
    rows = []
    for mangled, d in zip(out, demangled):
        d = d.replace("(anonymous namespace)::", "")
        short = d.split("(")[0].replace("void ", "").replace("dvid::", "")
        rows.append((short, out[mangled]))
    for short, c in sorted(rows, key=lambda t: t[0]):
        print("%-72s %6d " % (short[:72], c["_n"]) + " ".join("%7d" % c[k] for k in KEYS))
    tc = sorted({s for s, c in rows if c["UTCHMMA"]})
    hm = sorted({s for s, c in rows if c["HMMA"]})
    print("\nkernels issuing tcgen05.mma (UTCHMMA): " + ", ".join(tc))
    print("kernels issuing mma.sync (HMMA):       " + ", ".join(hm))


if __name__ == "__main__":
    main()
