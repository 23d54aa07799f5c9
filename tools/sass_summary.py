#!/usr/bin/env python
"""Per-kernel counts of the SASS mnemonics that prove the Blackwell paths (tcgen05 MMA, TMA loads / stores, TMEM loads, the legacy
mma.sync path) in the shipped library:  python tools/sass_summary.py > profiles/rNN_sass_summary.md   (cuobjdump, no GPU needed)"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "feed_forward_vqgan_clip_b200", "libffvc_sm100.so")
PAT = [("UTCHMMA", r"\bUTCHMMA"), ("UTCHMMA.2CTA", r"\bUTCHMMA\.2CTA"), ("UTMALDG", r"\bUTMALDG"), ("UTMALDG.2CTA", r"\bUTMALDG[.\w]*\.2CTA"),
       ("UTMASTG", r"\bUTMASTG"), ("LDTM", r"\bLDTM"), ("UTCBAR", r"\bUTCBAR"), ("SYNCS (mbarrier)", r"\bSYNCS"), ("HMMA.16816", r"\bHMMA\.16816"),
       ("LDSM", r"\bLDSM"), ("MUFU", r"\bMUFU"), ("RED/ATOM", r"\b(RED|ATOM|ATOMG|ATOMS)\b"), ("STG.E.ENL2.256", r"STG\.E\.ENL2\.256"),
       ("LDGSTS (cp.async)", r"\bLDGSTS")]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            kernels[cur]["_instr"] = 0
            continue
        if cur is None or "/*" not in line:
            continue
        if re.search(r"/\*[0-9a-f]{4,6}\*/", line):
            kernels[cur]["_instr"] += 1
            for name, pat in PAT:
                if re.search(pat, line):
                    kernels[cur][name] += 1
    demangle = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
    print("# SASS summary of libffvc_sm100.so (`cuobjdump -sass`, sm_100a): instruction counts per kernel\n")
    print("Blackwell-native paths: `UTCHMMA` = tcgen05.mma (`.2CTA` = cta_group::2), `UTMALDG` / `UTMASTG` = TMA tensor load / store, `LDTM` = tcgen05.ld (TMEM),")
    print("`UTCBAR` = tcgen05.commit, `SYNCS` = mbarrier ops; legacy warp path: `HMMA.16816` = mma.sync (attention kernels), `LDSM` = ldmatrix.\n")
    cols = [n for n, _ in PAT]
    print("| kernel | SASS instr. | " + " | ".join(cols) + " |")
    print("|---|---:|" + "---:|" * len(cols))
    tot = collections.Counter()
    for (mangled, c), name in zip(kernels.items(), demangle):
        short = re.sub(r"\(.*", "", name).replace("void ", "").replace("ffvc::", "")
        short = re.sub(r"\(int\)", "", short)
        if not any(c[n] for n in cols[:11]) and c["_instr"] < 400:
            continue                                   # small elementwise kernels without any of the marked instructions
        print("| `%s` | %d | " % (short[:90], c["_instr"]) + " | ".join(str(c[n]) if c[n] else "" for n in cols) + " |")
        tot.update(c)
    for mangled, c in kernels.items():
        pass
    alltot = collections.Counter()
    for c in kernels.values():
        alltot.update(c)
    print("| **whole library (%d kernels)** | %d | " % (len(kernels), alltot["_instr"]) + " | ".join(str(alltot[n]) for n in cols) + " |")


if __name__ == "__main__":
    main()
