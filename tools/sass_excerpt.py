"""Per-kernel counts and one sample line of the Blackwell-specific SASS instructions (tcgen05 MMA / TMEM load / TMA load & store /
mbarrier) in the built library -- evidence that the convolution kernels run on tcgen05 + TMA (no GPU needed)."""
import collections, os, re, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(root, "face_vae_b200", "libfacevae_b200.so")
MNEM = ("UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "UTMACMDFLUSH", "SYNCS", "UTCATOMSWS", "UBLKCP")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
cur, counts, sample = None, collections.OrderedDict(), {}
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.search(r"/\*[0-9a-f]{4}\*/\s+(.*?);", line)
    if not m:
        continue
    ins = m.group(1).strip()
    op = ins.split()[1] if ins.startswith("@") else ins.split()[0]
    base = op.split(".")[0]
    if base in MNEM:
        counts[cur][base] += 1
        sample.setdefault((cur, base), ins)
print("# tcgen05 / TMEM / TMA / mbarrier SASS per kernel of face_vae_b200/libfacevae_b200.so (cuobjdump -sass, sm_100a)\n")
print("UTCHMMA = tcgen05.mma (kind::f16), UTCBAR = tcgen05.commit -> mbarrier, LDTM = tcgen05.ld (TMEM -> registers), UTMALDG / UTMASTG = "
      "cp.async.bulk.tensor load / store (TMA), UTMAPF = TMA descriptor prefetch, SYNCS = mbarrier operations.\n")
for fn, c in counts.items():
    if not (c.get("UTCHMMA") or c.get("UTMALDG")):
        continue
    print(f"## {demangle(fn)[:150]}")
    print("   " + "  ".join(f"{k} x{v}" for k, v in sorted(c.items())))
    for k in ("UTCHMMA", "LDTM", "UTMALDG", "UTMASTG", "UTCBAR"):
        if (fn, k) in sample:
            print(f"     e.g. {sample[(fn, k)]}")
    print()
