"""`cuobjdump -sass` mnemonic counts of the built library: the evidence that the convolution path is tcgen05 / TMEM / TMA
(UTCHMMA, LDTM, UTCBAR, UTMALDG, UTMASTG) and which kernels use the legacy warp-level MMA (HMMA).

  python tools/sass_summary.py > profiles/rXX_sass_mnemonics.md
"""
import collections
import os
import re
import subprocess

LIB = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "seg2eye_b200", "_C", "libseg2eye_b200.so")
KEYS = [("UTCHMMA", "tcgen05.mma"), ("LDTM", "tcgen05.ld"), ("UTCBAR", "tcgen05.commit"), ("UTMALDG", "TMA load"), ("UTMASTG", "TMA store"),
        ("SYNCS", "mbarrier"), ("HMMA", "mma.sync"), ("REDG", "red.global"), ("LDGSTS", "cp.async")]
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
per = collections.OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        per[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m:
        op = m.group(1)
        for k, _ in KEYS:
            if op == k or op.startswith(k + "."):
                per[cur][k] += 1
names = subprocess.run(["c++filt"], input="\n".join(per), capture_output=True, text=True).stdout.splitlines()
tot = collections.Counter()
for c in per.values():
    tot.update(c)
print("# SASS evidence: `cuobjdump -sass seg2eye_b200/_C/libseg2eye_b200.so` (sm_100a), mnemonic counts per kernel\n")
print("Library totals: " + ", ".join("%s (%s) %d" % (k, d, tot[k]) for k, d in KEYS) + "\n")
print("Kernels with tcgen05 / TMA / warp-MMA instructions:\n")
for mangled, dem in zip(per, names):
    c = per[mangled]
    if not (c["UTCHMMA"] or c["UTMALDG"] or c["HMMA"] or c["LDTM"]):
        continue
    dem = re.sub(r"\(.*", "", dem.replace("(anonymous namespace)::", "").replace("void ", ""))
    print("- `%s`: %s" % (dem, ", ".join("%s %d" % (k, c[k]) for k, _ in KEYS if c[k])))
