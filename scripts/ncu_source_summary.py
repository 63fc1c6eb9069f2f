"""Summarise an `ncu --page source --csv` export: instruction mix by opcode (executed warp-instructions) and top stall sites."""
import csv, gzip, sys, collections
path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
f = gzip.open(path, "rt") if path.endswith(".gz") else open(path)
rows = list(csv.reader(f))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
ix = {h: i for i, h in enumerate(hdr)}
ops = collections.Counter(); samp = collections.Counter(); tot = 0; tots = 0
lines = []
for r in rows[hi + 1:]:
    if len(r) < len(hdr): continue
    src = r[ix["Source"]].strip()
    op = src.split()[0] if not src.startswith("@") else src.split()[1]
    op = op.split(".")[0]
    n = int(r[ix["Instructions Executed"]] or 0); s = int(r[ix["# Samples"]] or 0)
    ops[op] += n; samp[op] += s; tot += n; tots += s
    lines.append((s, n, src))
print("total executed warp-instructions %d, samples %d" % (tot, tots))
for op, n in ops.most_common(top):
    print("  %-10s %12d %5.1f%%   samples %5.1f%%" % (op, n, 100.0 * n / tot, 100.0 * samp[op] / max(tots, 1)))
print("top stall sites:")
for s, n, src in sorted(lines, reverse=True)[:top]:
    print("  %6d %10d  %s" % (s, n, src[:100]))
