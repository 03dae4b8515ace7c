"""Per-kernel SASS mnemonic counts that prove the Blackwell-native paths (tcgen05 = UTC*MMA, TMEM = LDTM/STTM,
TMA = UTMALDG, legacy tensor = HMMA): python tools/sass_evidence.py > profiles/rNN_sass_evidence.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "swift_b200", "libswift_b200.so")
PAT = re.compile(r"\b(UTC[A-Z]*MMA[.\w]*|UTMALDG[.\w]*|UTMASTG[.\w]*|UBLKCP[.\w]*|LDTM[.\w]*|STTM[.\w]*|UTCBAR[.\w]*|"
                 r"UTCATOMSWS[.\w]*|SYNCS\.[.\w]*|HMMA[.\w]*|LDSM[.\w]*|LDGSTS[.\w]*|ELECT|UCGABAR_\w+|MUFU\.EX2)")


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    kern = None
    counts = collections.OrderedDict()
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            kern = re.sub(r"\(.*", "", kern)
            counts[kern] = collections.Counter()
            continue
        if kern:
            for tok in PAT.findall(line):
                counts[kern][tok] += 1
    for k, c in counts.items():
        if not c:
            continue
        print(k)
        print("   " + "  ".join(f"{n}x{v}" for n, v in sorted(c.items())))


if __name__ == "__main__":
    main()
