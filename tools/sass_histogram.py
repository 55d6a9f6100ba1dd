"""Opcode evidence per object file of libddpmir (run here, no GPU): which kernels are Blackwell-native.

    python tools/sass_histogram.py > profiles/r2_sass_opcodes.txt

Counts, per build/*.o, the SASS mnemonics B200_PROFILING.md names: UTC*MMA (tcgen05.mma), LDTM/STTM (tcgen05.ld/st), UTMALDG /
UTMASTG / UBLKCP (TMA), HMMA (legacy mma.sync), LDGSTS (cp.async), LDSM (ldmatrix), MUFU.EX2."""
import os, re, subprocess, sys
from collections import Counter

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = os.path.join(ROOT, "ddpm_image_restoration_b200", "build")
PAT = re.compile(r"\b(UTC[A-Z]*MMA|LDTM|STTM|UTMALDG|UTMASTG|UBLKCP|HMMA\.[0-9]+|IMMA|LDGSTS|LDSM|MUFU\.EX2|SYNCS)\b")
print("# SASS opcode histogram per object (cuobjdump -sass), sm_100a build of", subprocess.run(
    ["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip())
for name in sorted(os.listdir(OBJ)):
    if not name.endswith(".o"):
        continue
    sass = subprocess.run(["cuobjdump", "-sass", os.path.join(OBJ, name)], capture_output=True, text=True).stdout
    c = Counter(m.group(1) for m in PAT.finditer(sass))
    kernels = len(re.findall(r"Function :", sass))
    line = "  ".join(f"{k} {v}" for k, v in sorted(c.items(), key=lambda kv: -kv[1]))
    print(f"{name:22s} {kernels:3d} kernels  {line if line else '-'}")
