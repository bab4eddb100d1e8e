"""Opcode histogram (stall samples + executed warp-instructions) of one kernel from a `--page source --csv` dump."""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[hi]; R = [r for r in rows[hi + 1:] if len(r) == len(h) and r[0] != "Address"]
ix = {n: i for i, n in enumerate(h)}
smp = [int(r[ix['# Samples']] or 0) for r in R]
exe = [int(r[ix['Instructions Executed']] or 0) for r in R]
src = [r[ix['Source']].strip() for r in R]
print("instr", len(R), "samples", sum(smp), "warp-instr", sum(exe))
ops = collections.Counter(); opx = collections.Counter()
for s, n, e in zip(src, smp, exe):
    t = s.split()
    op = t[0] if not t[0].startswith('@') else t[1]
    op = op.split('.')[0]
    ops[op] += n; opx[op] += e
for op, n in ops.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 25):
    print(f"{op:10s} samples {n:7d}  exec {opx[op]:9d}")
