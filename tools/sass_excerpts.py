"""profiles/*_sass_excerpts.txt: instruction histograms of the hot kernels and the K-b main loop with its control words decoded
   python tools/sass_excerpts.py > profiles/r2_sass_excerpts.txt     (cuobjdump only, no GPU)"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "poltergeist.jl_b200", "csrc", "libpoltergeist_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
funcs, cur = collections.OrderedDict(), None
for ln in txt.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        cur = m.group(1); funcs[cur] = []
    elif cur:
        funcs[cur].append(ln)


def instrs(lines):
    out, i = [], 0
    while i < len(lines):
        m = re.match(r"\s*/\*([0-9a-f]{4,5})\*/\s+(.*?);\s*/\* (0x[0-9a-f]+) \*/", lines[i])
        if m and i + 1 < len(lines):
            m2 = re.match(r"\s*/\* (0x[0-9a-f]+) \*/", lines[i + 1])
            hi = int(m2.group(1), 16) if m2 else 0
            out.append((int(m.group(1), 16), m.group(2).strip(), (hi >> 41) & 0xf, (hi >> 46) & 7, (hi >> 49) & 7, (hi >> 52) & 0x3f))
            i += 2
        else:
            i += 1
    return out


def opname(t):
    t = [w for w in t.split() if not w.startswith("@")]
    return t[0] if t else "?"


print("SASS of libpoltergeist_b200.so (cuobjdump -sass, sm_100a), round 2 final: instruction histograms of the hot kernels and excerpts.")
print("No UTMALDG / UTMASTG / UTC*MMA / LDTM / STTM anywhere: the path is FP64 (tcgen05 has no FP64 kind); the tensor-core")
print("instruction it does use is DMMA.8x8x4 (FP64 mma.sync m8n8k4) in basis_mma_kernel.  multimem.st.relaxed.sys (transform_kernel's")
print("multi-GPU epilogue) lowers to STG.E.64/128.STRONG.SYS on the multicast address: the replication is done by the NVSwitch, not the SM.")
print("mbarrier arrive / try_wait (basis_mma_kernel's split barrier) are the SYNCS.* instructions; LDG.E.NA = ld.global.nc.L1::no_allocate.\n")
whole = collections.Counter()
hot = [k for k in funcs if any(s in k for s in ("basis_mma_kernelILi32ELi8ELb0", "transform_kernelILi12", "basis_kernelILi16ELi8ELi4ELb0", "invert_branches_kernelILi5ELb0"))]
for k, lines in funcs.items():
    ins = instrs(lines)
    for _, t, *_ in ins:
        whole[opname(t).split(".")[0]] += 1
    if k in hot:
        c = collections.Counter(opname(t).split(".")[0] for _, t, *_ in ins)
        full = collections.Counter(opname(t) for _, t, *_ in ins)
        print("== %s: %d instructions" % (k, len(ins)))
        print("   " + ", ".join("%s %d" % kv for kv in c.most_common(18)))
        note = [(n, v) for n, v in sorted(full.items()) if re.match(r"(DMMA|LDG\.E\.(NA\.)?128|LDS\.128|STS\.128|STG\.E\.(64|128)(\.STRONG\.SYS)?$|SYNCS|LDL|STL|ATOMG)", n)]
        print("   of note: " + ", ".join("%s x%d" % nv for nv in note))
print("\nwhole library: " + ", ".join("%s %d" % (n, whole[n]) for n in ("DMMA", "DFMA", "DADD", "DMUL", "LDS", "STS", "LDG", "STG", "LDL", "STL", "UTMALDG", "UTMASTG", "UTCHMMA", "LDTM", "STTM")))
# K-b main loop: from the first DMMA of the loop to the mbarrier arrive, control words decoded
k = [k for k in hot if "basis_mma" in k][0]
ins = instrs(funcs[k])
first = next(i for i, x in enumerate(ins) if "DMMA" in x[1])
last = next(i for i, x in enumerate(ins) if "SYNCS.ARRIVE" in x[1])
print("\n== %s: one chunk of the main loop (first MMA ... mbarrier arrive), control words decoded:" % k)
print("   st = stall count, wb/rb = write/read scoreboard set (7 = none), wm = scoreboards waited for.  Every DMMA carries st=15 (+ a NOP): the")
print("   warp itself waits out the pipe's 16-cycle issue interval; the flush (LDS ... STG.E.128) sits BEHIND each burst of 16 MMAs.")
for a, t, st, wb, rb, wm in ins[first - 6:last + 2]:
    print("   %05x st=%2d wb=%d rb=%d wm=%s  %s" % (a, st, wb, rb, format(wm, "06b"), t[:100]))
k = [k for k in hot if "transform_kernelILi12" in k][0]
print("\n== %s: the multicast epilogue (multimem.st) and the streaming column loads" % k)
for a, t, st, wb, rb, wm in instrs(funcs[k]):
    if "STRONG.SYS" in t or "LDG.E.NA" in t:
        print("   %05x  %s" % (a, t[:110]))
