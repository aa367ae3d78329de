"""Per-kernel counts of the SASS mnemonics that show the Blackwell paths (tcgen05 MMA = UTC*MMA, TMEM loads/stores =
LDTM/STTM, TMA = UTMALDG/UTMASTG/UBLKCP, legacy tensor path = HMMA) in the built library.

    python scripts/sass_evidence.py [lib.so] > profiles/r01_sass_evidence.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(
    ROOT, "multimodal-sensor-fusion-with-attention-rajeevatla_b200", "libmsf_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
PAT = [("UTC*MMA", re.compile(r"\bUTC\w*MMA\b")), ("LDTM", re.compile(r"\bLDTM\b")), ("STTM", re.compile(r"\bSTTM\b")),
       ("UTMALDG", re.compile(r"\bUTMALDG\b")), ("UTMASTG", re.compile(r"\bUTMASTG\b")),
       ("UBLKCP", re.compile(r"\bUBLKCP\b")), ("LDGMC", re.compile(r"\bLDGMC\b")), ("HMMA", re.compile(r"\bHMMA\b")), ("RED/ATOM", re.compile(r"\b(REDG?|REDUX|ATOMG?|ATOMS)\b"))]
counts = collections.OrderedDict()
name = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = name.replace("(anonymous namespace)::", "").replace("msf::", "")
        name = re.sub(r"\(.*", "", name).replace("void ", "")
        if name in counts:   # same kernel name from another translation unit
            name += " [2]"
        counts[name] = collections.Counter()
        continue
    if name is None:
        continue
    counts[name]["instructions"] += 1 if re.match(r"\s+/\*[0-9a-f]{4,5}\*/\s+\S", line) else 0
    for key, pat in PAT:
        if pat.search(line):
            counts[name][key] += 1
cols = ["instructions"] + [k for k, _ in PAT]
print("# cuobjdump -sass of %s (sm_100a): per-kernel mnemonic counts" % os.path.basename(lib))
print("%-46s" % "kernel" + "".join("%13s" % c for c in cols))
for k, c in counts.items():
    print("%-46s" % k[:45] + "".join("%13d" % c[x] for x in cols))
