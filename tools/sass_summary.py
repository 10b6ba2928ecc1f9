"""Instruction-class summary of the shipped library's SASS, per kernel (cuobjdump -sass): the mnemonics that prove the
Blackwell path - UTCHMMA (tcgen05.mma), UTMALDG / UTMASTG (TMA load / store), LDTM (tcgen05.ld), UTCBAR (tcgen05.commit),
SYNCS (mbarrier), REDUX - next to the classic pipes.      python tools/sass_summary.py > profiles/r2_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "object-detection-yolov3_b200", "libyolo3_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
classes = [("UTCHMMA.2CTA", r"\bUTCHMMA\.2CTA"), ("UTCHMMA", r"\bUTCHMMA\b(?!\.2CTA)"), ("UTMALDG.IM2COL", r"\bUTMALDG\.\dD\.IM2COL"),
           ("UTMALDG", r"\bUTMALDG\b"), ("UTMASTG", r"\bUTMASTG\b"), ("LDTM", r"\bLDTM\b"), ("UTCBAR", r"\bUTCBAR\b"),
           ("SYNCS", r"\bSYNCS\b"), ("REDUX", r"\bREDUX\b"), ("MATCH", r"\bMATCH\b"), ("ATOM/RED", r"\b(ATOMG|ATOMS|ATOM|RED)\b"),
           ("FFMA2/FMUL2", r"\b(FFMA2|FMUL2|FADD2)\b"), ("HMMA", r"\bHMMA\b"), ("LDG", r"\bLDG\b"), ("STG", r"\bSTG\b"),
           ("LDS", r"\bLDS\b"), ("STS", r"\bSTS\b"), ("generic LD/ST", r"\b(LD|ST)\.E\b"), ("LDL/STL", r"\b(LDL|STL)\b"), ("MUFU", r"\bMUFU\b")]
cur, rows, counts, total = None, [], None, 0
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        if cur:
            rows.append((cur, total, counts))
        cur, counts, total = m.group(1), collections.Counter(), 0
        continue
    if cur and re.match(r"\s+/\*[0-9a-f]{4}\*/", line):
        total += 1
        for name, pat in classes:
            if re.search(pat, line):
                counts[name] += 1
if cur:
    rows.append((cur, total, counts))
dem = subprocess.run(["cu++filt"] + [r[0] for r in rows], capture_output=True, text=True).stdout.splitlines()
print("# %s - sm_100a SASS, instructions per kernel by class (only kernels of this library)" % os.path.basename(so))
print("%-78s %7s  %s" % ("kernel", "instr", "classes"))
for (name, tot, c), d in sorted(zip(rows, dem), key=lambda t: -t[0][1]):
    short = re.sub(r"\(.*", "", d).replace("y3::", "")
    print("%-78s %7d  %s" % (short[:78], tot, " ".join("%s=%d" % (k, c[k]) for k, _ in classes if c[k])))
