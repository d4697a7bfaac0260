"""Per-kernel count of the SASS mnemonics that prove Blackwell-native code paths (profiles/sass_summary.txt)."""
import collections, re, subprocess, sys
so = sys.argv[1] if len(sys.argv) > 1 else "neuronx_distributed_b200/_build/nxd_b200_C.so"
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
pat = re.compile(r"\b(LDGMC\S*|STGMC\S*|REDGMC\S*|REDG\S*|UTCHMMA(?:\.2CTA)?|UTCQMMA\S*|UTMALDG\S*|UTMASTG\S*|UBLK\S*|LDTM\S*|STTM\S*|UTCBAR\S*|UTCCP\S*|SYNCS\.\S+|HMMA\S*|MEMBAR\.\S+|RED\.E\S*|ATOM\S*|(?:LD|ST)\.E\S*SYS\S*|LDG\.E\S*|STG\.E\S*)")
cur, counts = None, collections.OrderedDict()
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()[:150]
        counts[cur] = collections.Counter()
        continue
    if cur:
        for mn in pat.findall(line):
            counts[cur][mn.split("(")[0]] += 1
for k, c in counts.items():
    if not c:
        continue
    print(k)
    print("   " + "  ".join(f"{m}:{n}" for m, n in sorted(c.items())))
