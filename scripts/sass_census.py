#!/usr/bin/env python
"""SASS mnemonic census of pykrige_b200/libkrige_b200.so (cuobjdump -sass; needs no GPU).

    python scripts/sass_census.py > profiles/r02/sass_census_final_build.md

Counts, per kernel family, the instructions that prove which hardware path a kernel uses
(B200_PROFILING.md: DMMA = fp64 tensor pipe, UTCHMMA/UTCIMMA = tcgen05.mma kind::tf32 / kind::i8,
LDTM = tcgen05.ld, UBLKCP = cp.async.bulk through the TMA engine, UTMALDG = tensor-map TMA loads,
SYNCS = mbarrier traffic, UTCBAR = tcgen05.commit)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "pykrige_b200", "libkrige_b200.so")
WATCH = ["DMMA", "UTCHMMA", "UTCIMMA", "UTCBAR", "UTCATOMSWS", "LDTM", "UBLKCP", "UTMALDG", "UTMASTG", "SYNCS",
         "DFMA", "DMUL", "DADD", "MUFU", "LDGSTS", "REDUX", "SHFL", "BAR", "ATOMG", "STL", "LDL"]


def demangled_family(name):
    try:
        full = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
    except OSError:
        full = name
    m = re.match(r"(?:void\s+)?([A-Za-z_][A-Za-z0-9_:]*)", full)
    return m.group(1) if m else full


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    per_fn = collections.defaultdict(collections.Counter)
    fn = None
    for ln in sass.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            fn = m.group(1)
            continue
        m = re.search(r"/\*[0-9a-f]+\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_]*)", ln)
        if m and fn:
            per_fn[fn][m.group(1)] += 1
    fam = collections.defaultdict(lambda: [0, collections.Counter()])
    for f, c in per_fn.items():
        e = fam[demangled_family(f)]
        e[0] += 1
        e[1].update(c)
    cols = [w for w in WATCH if any(e[1][w] for e in fam.values())]
    print("# SASS census of `pykrige_b200/libkrige_b200.so` (sm_100a, `cuobjdump -sass`, `scripts/sass_census.py`)\n")
    print("%d kernels (template instantiations) in %d families. Counts are static instruction counts summed over "
          "the instantiations of a family.\n" % (len(per_fn), len(fam)))
    print("| kernel family | inst. | total SASS | " + " | ".join(cols) + " |")
    print("|---|---|---|" + "---|" * len(cols))
    for name in sorted(fam, key=lambda k: -sum(fam[k][1].values())):
        n, c = fam[name]
        print("| `%s` | %d | %d | %s |" % (name, n, sum(c.values()), " | ".join(str(c[w]) if c[w] else "" for w in cols)))
    tot = collections.Counter()
    for e in fam.values():
        tot.update(e[1])
    print("| **library** | %d | %d | %s |" % (len(per_fn), sum(tot.values()), " | ".join(str(tot[w]) for w in cols)))
    print("\nNo `HMMA`/`IMMA`/`QMMA` (legacy mma.sync low-precision) and no `UTMALDG`/`UTMASTG`: the operand tiles are "
          "stored operand-ready, so\nplain 1-D bulk copies (`UBLKCP`) feed both the DMMA and the tcgen05 kernels "
          "(DESIGN.md §4)." if not (tot["HMMA"] or tot["IMMA"] or tot["UTMALDG"]) else "")
    return 0


if __name__ == "__main__":
    sys.exit(main())
