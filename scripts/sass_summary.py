"""Blackwell-native evidence, from the shipped binary: per kernel of dafne_b200/libdafne_b200.so the number of
tcgen05 / TMEM / TMA instructions in its SASS (cuobjdump -sass), written to profiles/sass_summary.md.

  python scripts/sass_summary.py [path/to/lib.so] [out.md]

SASS mnemonics (B200_PROFILING.md): tcgen05.mma -> UTC*MMA, tcgen05.ld/st -> LDTM/STTM, cp.async.bulk.tensor ->
UTMALDG/UTMASTG, tcgen05.commit -> UTCBAR, tcgen05.alloc -> UTCATOMSWS; HMMA would be the legacy mma.sync path.
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ["UTCHMMA", "UTCBAR", "UTCATOMSWS", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "SYNCS", "HMMA", "LDL", "STL"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def main():
    lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "dafne_b200", "libdafne_b200.so")
    dst = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "profiles", "sass_summary.md")
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    arch = sorted(set(re.findall(r"arch = (sm_\w+)", sass)))
    per = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            per[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m and cur:
            per[cur]["_total"] += 1
            op = m.group(1)
            for k in KEYS:
                if op.startswith(k):
                    per[cur][k] += 1
    names = demangle(list(per))
    rows = []
    for fn, c in per.items():
        nm = re.sub(r"\(.*", "", names[fn]).replace("void ", "").replace("dafne::", "")
        rows.append((nm, c))
    rows.sort(key=lambda r: (-r[1]["UTCHMMA"], r[0]))
    tot = collections.Counter()
    for _, c in rows:
        tot.update(c)
    with open(dst, "w") as f:
        f.write(f"# SASS summary of `{os.path.relpath(lib, ROOT)}` ({', '.join(arch)}; `cuobjdump -sass`, scripts/sass_summary.py)\n\n")
        f.write("Instruction counts per kernel (static). `UTCHMMA` = tcgen05.mma kind::f16, `LDTM` = tcgen05.ld, "
                "`UTMALDG` / `UTMASTG` = TMA tensor load / store, `UTCBAR` = tcgen05.commit, `UTCATOMSWS` = TMEM "
                "alloc / dealloc, `SYNCS` = mbarrier ops. `HMMA` (legacy mma.sync) must be 0 everywhere; `LDL` / `STL` = "
                "local-memory (spill / stack) accesses.\n\n")
        f.write("| kernel | instructions | " + " | ".join(KEYS) + " |\n|---|---|" + "---|" * len(KEYS) + "\n")
        f.write(f"| **all {len(rows)} kernels** | {tot['_total']} | " + " | ".join(str(tot[k]) for k in KEYS) + " |\n")
        for nm, c in rows:
            f.write(f"| `{nm}` | {c['_total']} | " + " | ".join(str(c[k]) if c[k] else "" for k in KEYS) + " |\n")
    print(open(dst).read()[:2500])
    assert tot["HMMA"] == 0, "legacy tensor-core instructions found"
    assert tot["UTCHMMA"] > 0 and tot["UTMALDG"] > 0 and tot["LDTM"] > 0


if __name__ == "__main__":
    main()
