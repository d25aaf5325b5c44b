"""profiles/r02_sass_tc.txt: per-kernel counts of the Blackwell-native SASS mnemonics in libbpmult_b200.so (cuobjdump -sass):
UTCHMMA (tcgen05.mma; .2CTA = cta_group::2), LDTM / STTM (tcgen05.ld / st), UTMALDG / UTMASTG / UTMAREDG (TMA tensor load / store /
reduce), UBLKCP (1-D bulk copy), HMMA (legacy mma.sync: must be 0).  Usage: python scripts/sass_summary.py > profiles/r02_sass_tc.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "bpmult_b200", "libbpmult_b200.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
dem = {}
pat = ("UTCHMMA.2CTA", "UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UBLKCP", "HMMA", "MUFU.EX2")
cur, counts = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        counts[cur]["_instr"] = 0
        continue
    if cur is None or "/*" not in line:
        continue
    body = line.split("/*")[1] if line.strip().startswith("/*") else line
    if re.search(r"/\*[0-9a-f]{4}\*/", line):
        counts[cur]["_instr"] += 1
        for p in pat:
            if re.search(r"\b" + re.escape(p) + r"\b", line) if p != "UTCHMMA" else re.search(r"\bUTCHMMA\b(?!\.2CTA)", line):
                counts[cur][p] += 1
names = subprocess.run(["c++filt"], input="\n".join(counts.keys()), capture_output=True, text=True).stdout.splitlines()
rev = subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
print("# cuobjdump -sass bpmult_b200/libbpmult_b200.so  (commit %s): kernels that carry tensor-core / TMEM / TMA instructions" % rev)
print("%-110s %7s %s" % ("kernel", "SASS", "  ".join("%s" % p for p in pat)))
tot = collections.Counter()
for (mangled, c), name in zip(counts.items(), names):
    if not any(c[p] for p in pat if p != "MUFU.EX2"):
        continue
    short = re.sub(r"\(.*", "", name)[:108]
    print("%-110s %7d %s" % (short, c["_instr"], "  ".join("%*d" % (len(p), c[p]) for p in pat)))
    tot.update({p: c[p] for p in pat})
print("%-110s %7s %s" % ("TOTAL", "", "  ".join("%*d" % (len(p), tot[p]) for p in pat)))
assert tot["HMMA"] == 0, "legacy mma.sync found"
