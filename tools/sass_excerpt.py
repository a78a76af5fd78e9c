"""Per kernel, the SASS mnemonics that show what hardware path it uses (tcgen05 / TMEM / TMA / dot-product / SIMD-in-register video
instructions). Usage: python tools/sass_excerpt.py > profiles/rNN_sass_excerpt.txt"""
import re
import subprocess
import sys
from collections import Counter, OrderedDict

so = sys.argv[1] if len(sys.argv) > 1 else "multiagent_orb_slam2_b200/lib/liborb_b200.so"
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
names = {}
keep = re.compile(r"^(UTC|LDTM|STTM|UTMA|UBLKCP|UCGABAR|IDP|POPC|VABSDIFF|VIMNMX3|SYNCS\.ARRIVE\.TRANS64\.RED)")
per = OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        per[cur] = Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Za-z0-9_.]+)", line)
    if m and cur and keep.match(m.group(1)):
        per[cur][m.group(1)] += 1
dem = subprocess.run(["c++filt"] + list(per), capture_output=True, text=True).stdout.splitlines()
print("SASS mnemonics of %s (cuobjdump -sass, sm_100a), per kernel: instruction x count\n" % so)
for mangled, d in zip(per, dem):
    if per[mangled]:
        print("%-60s %s" % (d.split("(")[0], "  ".join("%s x%d" % kv for kv in sorted(per[mangled].items()))))
