"""profiles/rNN_sass_summary.txt: counts of the SASS mnemonics that prove tcgen05 / TMEM / TMA / clusters in the in-tree
library (`cuobjdump -sass`), in total and per kernel.

    python tools/sass_summary.py > profiles/r02_sass_summary.txt
"""
import collections
import datetime
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "highway-rope-ppo_b200", "lib", "libhrp_b200.so")
MNEMONICS = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "UTCATOM", "SYNCS", "UCGABAR",
             "MUFU.EX2", "MUFU.RSQ", "DFMA", "REDUX", "FFMA2", "HMMA", "IMMA"]

sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
total = collections.Counter()
per = collections.OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        per[cur] = collections.Counter()
        continue
    if cur is None or not re.match(r"\s+/\*[0-9a-f]{4,}\*/", line):   # instruction lines carry their address
        continue
    for mn in MNEMONICS:
        if re.search(r"\b" + re.escape(mn), line):
            total[mn] += 1
            per[cur][mn] += 1
    per[cur]["_all"] += 1

print(f"# SASS summary of highway-rope-ppo_b200/lib/libhrp_b200.so (cuobjdump -sass, sm_100a), {datetime.date.today()}")
print("# command: python tools/sass_summary.py   (counts lines of `cuobjdump -sass` that contain the mnemonic)")
for mn in MNEMONICS:
    print(f"{mn:12s} {total[mn]}")
print("\n# per kernel: instructions, then the non-zero counts of the mnemonics above")
for name, c in per.items():
    short = subprocess.run(["c++filt", "-p", name], capture_output=True, text=True).stdout.strip() or name
    short = re.sub(r"\(anonymous namespace\)::", "", short)
    tags = ", ".join(f"{k} {v}" for k, v in c.items() if k != "_all")
    print(f"{short[:110]}: {c['_all']} instructions" + (f"; {tags}" if tags else ""))
