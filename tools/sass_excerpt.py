#!/usr/bin/env python
"""Per-kernel counts of the SASS mnemonics that show how data moves (TMA, mbarrier, tcgen05, vector reductions, packed FMA), plus the
first occurrences in context, from the built library.  usage: python tools/sass_excerpt.py > profiles/rNN_sass_excerpts.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "combo-avs_b200", "libmsda_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
demangle = lambda n: subprocess.run(["cu++filt", n], capture_output=True, text=True).stdout.strip() or n
PAT = re.compile(r"\b(UTMALDG[.\w]*|UTMASTG[.\w]*|UBLKCP[.\w]*|SYNCS[.\w]*|UTCMMA[.\w]*|UTCHMMA[.\w]*|UTCBAR[.\w]*|UTCCP[.\w]*|LDTM[.\w]*|STTM[.\w]*|"
                 r"LDGSTS[.\w]*|RED\.E\.ADD\.F32[.\w]*|REDG[.\w]*|FFMA2|ATOMS[.\w]*|LDS\.128|SHFL[.\w]*)")
counts, ctx, cur, lines = collections.OrderedDict(), {}, None, []
for ln in sass.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        ctx[cur] = []
        continue
    if cur is None:
        continue
    m = PAT.search(ln)
    if m:
        counts[cur][m.group(1)] += 1
        if m.group(1).startswith(("UTMA", "SYNCS", "UTC", "LDTM")) and len(ctx[cur]) < 6:
            ctx[cur].append(re.sub(r"\s*/\* 0x[0-9a-f]+ \*/\s*$", "", ln).strip())
want = [k for k in counts if any(t in k for t in ("slab32_tma_kernelILi4E", "bwd_sorted32_kernelIffLi4ELi8E", "bwd_sorted32_kernelIffLi4ELi12E", "linear"))]
print("# cuobjdump -sass combo-avs_b200/libmsda_b200.so (sm_100a): mnemonic counts of the default S4 kernels, then all kernels that use TMA / tcgen05")
for k in want:
    print("\n## %s" % demangle(k)[:200])
    print("   " + ", ".join("%s x%d" % kv for kv in sorted(counts[k].items())))
    for c in ctx[k]:
        print("      " + c)
print("\n## kernels in the library: %d; with UTMALDG: %d; with UTCHMMA/UTCMMA (tcgen05.mma): %d; with LDGSTS (cp.async): %d" % (
    len(counts), sum(1 for c in counts.values() if any(k.startswith("UTMALDG") for k in c)),
    sum(1 for c in counts.values() if any(k.startswith(("UTCHMMA", "UTCMMA")) for k in c)),
    sum(1 for c in counts.values() if any(k.startswith("LDGSTS") for k in c))))
