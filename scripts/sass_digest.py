"""Per-kernel SASS opcode digest of libdig_b200.so: counts of the Blackwell-native mnemonics (tcgen05.mma -> UTC*MMA, tcgen05.ld/st ->
LDTM/STTM, TMA -> UTMALDG/UTMASTG/UTMAREDG/UTMAPF, packed fp32 -> FFMA2) next to the legacy tensor path (HMMA), per kernel.
Runs without a GPU (cuobjdump):  python scripts/sass_digest.py > profiles/r2_sass_digest.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "dig_b200", "libdig_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
dem = subprocess.run(["cu++filt"], input="\n".join(re.findall(r"Function : (\S+)", out)), capture_output=True, text=True).stdout.splitlines()
names = dict(zip(re.findall(r"Function : (\S+)", out), dem))
KEYS = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UTMAPF", "UTCBAR", "SYNCS", "FFMA2", "MUFU", "HMMA", "LDG", "STG",
        "RED", "ATOM"]
cur, rows = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = names.get(m.group(1), m.group(1))
        rows[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        op = m.group(1)
        base = op.split(".")[0]
        rows[cur][base] += 1
        if base == "UTCHMMA" and ".2CTA" in op:
            rows[cur]["UTCHMMA.2CTA"] += 1
        rows[cur]["_total"] += 1
print("SASS opcode digest of %s (sm_100a)" % os.path.relpath(lib, ROOT))
print("columns: " + " ".join(KEYS) + " | total instructions")


def short(n):
    n = n.replace("void ", "").replace("(bool)", "").replace("(int)", "")
    n = re.sub(r">\(.*", ">", n) if ">(" in n else re.sub(r"\(.*", "", n)
    return n[:78]


for n, c in rows.items():
    if c["_total"] == 0:
        continue
    print("%-80s %s | %d" % (short(n), " ".join("%5d" % c[k] for k in KEYS), c["_total"]))
tot = collections.Counter()
for c in rows.values():
    tot.update(c)
print("%-80s %s | %d" % ("ALL KERNELS", " ".join("%5d" % tot[k] for k in KEYS), tot["_total"]))
print("legacy mma.sync (HMMA) instructions: %d; wgmma (HGMMA): %d" % (tot["HMMA"], tot["HGMMA"]))
