"""Summarise an `ncu --page source --csv` export: instruction mix by opcode (executed warp-instructions) and top stall sites.
An export holds one section per captured launch ("Kernel Name" row, "Address,..." header row, one row per instruction); captures
taken with `-c N` hold N sections, which are summed here."""
import collections
import csv
import gzip
import sys

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
f = gzip.open(path, "rt") if path.endswith(".gz") else open(path)
ops = collections.Counter(); samp = collections.Counter(); tot = 0; tots = 0
lines = collections.OrderedDict()
ix, sections = None, 0


def num(v):
    try:
        return int(float(v))
    except ValueError:
        return 0


for r in csv.reader(f):
    if not r:
        continue
    if r[0] == "Kernel Name":
        ix = None
        continue
    if r[0] == "Address":
        ix = {h: i for i, h in enumerate(r)}
        sections += 1
        continue
    if ix is None or len(r) <= max(ix["Source"], ix["Instructions Executed"], ix["# Samples"]):
        continue
    src = r[ix["Source"]].strip()
    if not src:
        continue
    toks = src.split()
    op = toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]
    op = op.split(".")[0]
    n, s = num(r[ix["Instructions Executed"]]), num(r[ix["# Samples"]])
    ops[op] += n; samp[op] += s; tot += n; tots += s
    key = (r[ix["Address"]][-6:], src)
    a = lines.get(key, (0, 0))
    lines[key] = (a[0] + s, a[1] + n)
print("launches captured %d; total executed warp-instructions %d, samples %d" % (sections, tot, tots))
for op, n in ops.most_common(top):
    print("  %-10s %12d %5.1f%%   samples %5.1f%%" % (op, n, 100.0 * n / max(tot, 1), 100.0 * samp[op] / max(tots, 1)))
print("top stall sites:")
for (addr, src), (s, n) in sorted(lines.items(), key=lambda kv: -kv[1][0])[:top]:
    print("  %6d %10d  %s" % (s, n, src[:100]))
