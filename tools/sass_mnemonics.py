"""Count, per kernel of libsimx_b200.so, the SASS mnemonics that prove the Blackwell-native paths
(cuobjdump -sass simxns_b200/libsimx_b200.so | python tools/sass_mnemonics.py > profiles/r01_sass_mnemonics.txt)."""
import collections
import re
import subprocess
import sys

WANT = ["UTCHMMA", "UTCBAR", "UTMALDG", "UTMASTG", "UTMAREDG", "LDTM", "STTM", "SYNCS", "HMMA", "REDG", "RED.", "ATOMS",
        "MUFU.TANH", "MUFU.EX2", "MATCH.ANY", "FFMA"]
cur, counts = None, collections.OrderedDict()
for line in sys.stdin:
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.search(r"/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]+)", line)
    if m:
        op = m.group(1)
        for w in WANT:
            if op.startswith(w):
                counts[cur][w] += 1
names = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
print("cuobjdump -sass simxns_b200/libsimx_b200.so (sm_100a): counts of the mnemonics that prove the Blackwell-native paths")
print("UTCHMMA = tcgen05.mma, UTCBAR = tcgen05.commit, LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG/UTMAREDG = TMA load / store /"
      " reduce-add, SYNCS = mbarrier ops, HMMA = legacy mma.sync (attention.cu only: forward for 256 < S <= 512), FFMA-only ="
      " the fp32-class reference attention\n")
for fn, name in zip(counts, names):
    c = counts[fn]
    keep = {k: v for k, v in c.items() if k != "FFMA" or not any(x in c for x in ("UTCHMMA", "HMMA"))}
    if not keep:
        continue
    name = re.sub(r"\(.*", "", name).replace("void simx::", "").replace("simx::", "")[:70]
    print(f"{name:70s} " + "  ".join(f"{k}={v}" for k, v in keep.items()))
