#!/bin/bash
# Per-kernel counts of the SASS mnemonics that prove the Blackwell-native path (tcgen05 MMA / TMEM traffic / TMA / clusters)
# in the built library: bash scripts/sass_extract.sh > profiles/<tag>_sass_tc.txt
LIB=${1:-c-attl3_b200/libcattl3_b200.so}
echo "# cuobjdump -sass $LIB: instruction counts per kernel (sm_100a)"
cuobjdump -sass $LIB | python3 -c '
import sys, re, collections, subprocess
ops = ("UTCHMMA", "UTMALDG", "UTCBAR", "STTM", "LDTM", "UCGABAR", "DMMA", "RED", "LDGSTS", "HMMA", "SYNCS")
cur, tab = None, collections.OrderedDict()
for line in sys.stdin:
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1); tab[cur] = collections.Counter(); continue
    m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        full = m.group(1); base = full.split(".")[0]
        if base in ops:
            tab[cur][base + (".2CTA" if ".2CTA" in full else "")] += 1
names = list(tab)
dem = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
for n, d in sorted(zip(names, dem), key=lambda t: t[1]):
    if tab[n]:
        print("%-110s %s" % (d[:110], " ".join("%s=%d" % kv for kv in sorted(tab[n].items()))))
'
