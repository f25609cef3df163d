"""SASS mnemonic histogram of libnib.so (cuobjdump -sass): per kernel, the counts of the instructions that prove the
tcgen05 / TMEM / TMA / DMMA paths are what was compiled (B200_PROFILING.md), plus the overall top mnemonics.
usage: python tools/sass_histogram.py [libnib.so] > profiles/r02_sass_histogram.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "network_interpretation_imagenet_b200", "libnib.so")
KEY = ("UTCHMMA", "UTCQMMA", "UTCBAR", "UTMALDG", "UTMASTG", "UTMAPF", "UTMACMDFLUSH", "LDTM", "STTM", "UTCATOMSWS", "SYNCS", "DMMA", "HMMA",
       "ELECT", "UCGABAR_ARV", "ACQBULK")
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
per, total = collections.OrderedDict(), collections.Counter()
cur = None
for line in txt.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
        cur = re.sub(r"\(.*", "", cur).replace("void ", "").replace("nib::", "")
        per[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*(?:\.[A-Z0-9_.]+)?)", line)
    if m and cur:
        full = m.group(1)
        base = full.split(".")[0]
        per[cur][full if base in KEY else base] += 1
        total[base] += 1
print(f"# {os.path.relpath(lib, ROOT)}: {len(per)} kernels, {sum(total.values())} SASS instructions (sm_100a)")
print("\n## tcgen05 / TMEM / TMA / DMMA instructions per kernel (full mnemonic with modifiers)")
for k, c in per.items():
    hits = {m: n for m, n in c.items() if m.split(".")[0] in KEY}
    if hits:
        print(f"\n{k}  [{sum(c.values())} instructions]")
        for m, n in sorted(hits.items(), key=lambda t: (-t[1], t[0])):
            print(f"    {n:6d}  {m}")
print("\n## whole library, top 40 base mnemonics")
for m, n in total.most_common(40):
    print(f"    {n:7d}  {m}")
print("\n## whole library, key families")
for k in KEY:
    if total[k]:
        print(f"    {total[k]:7d}  {k}")
