"""Blackwell-native evidence from the built objects: counts of tcgen05 / TMEM / TMA SASS mnemonics per object file
(cuobjdump -sass of efficientq_b200/build/*.o, sm_100a).  python tools/sass_table.py > profiles/r02_sass.md"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "efficientq_b200", "build")
MNEMONICS = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "UTMALDG", "UBLKCP", "SYNCS", "DFMA", "HMMA"]
print("# SASS mnemonic counts per object (`cuobjdump -sass`, sm_100a)\n")
print("UTCHMMA = tcgen05.mma kind::f16, UTCQMMA = kind::f8f6f4, LDTM = tcgen05.ld (TMEM), UTMALDG = TMA tensor load, "
      "UBLKCP = 1-D bulk copy, SYNCS = mbarrier, DFMA = fp64 FMA, HMMA = legacy mma.sync (none expected)\n")
print("| object | " + " | ".join(MNEMONICS) + " | kernels |")
print("|---|" + "---|" * (len(MNEMONICS) + 1))
for f in sorted(os.listdir(BUILD)):
    if not f.endswith(".o"):
        continue
    sass = subprocess.run(["cuobjdump", "-sass", os.path.join(BUILD, f)], stdout=subprocess.PIPE, text=True).stdout
    counts = [len(re.findall(r"\b" + m + r"\b", sass)) for m in MNEMONICS]
    kernels = len(re.findall(r"Function : ", sass))
    print(f"| {f} | " + " | ".join(str(c) for c in counts) + f" | {kernels} |")
