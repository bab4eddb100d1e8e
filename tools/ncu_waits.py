"""Attribute warp-stall samples of a kernel in an .ncu-rep to mbarrier waits / instruction classes."""
import csv, subprocess, sys, io, collections
rep, skip = sys.argv[1], int(sys.argv[2])
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
blocks = []
for r in csv.reader(io.StringIO(out)):
    if r and r[0] == "Kernel Name":
        blocks.append([r])
    elif blocks:
        blocks[-1].append(r)
rows = blocks[skip]
name = rows[0][1]; h = rows[1]; R = [r for r in rows[2:] if len(r) == len(h)]
ix = {n: i for i, n in enumerate(h)}
smp = [int(r[ix['# Samples']] or 0) for r in R]
src = [r[ix['Source']].strip() for r in R]
exe = [int(r[ix['Instructions Executed']] or 0) for r in R]
print(name[:100]); print("total samples", sum(smp), "total warp-instructions", sum(exe))
stalls = [n for n in h if n.startswith('stall_') and 'Not Issued' not in n]
agg = collections.Counter()
for r in R:
    for n in stalls: agg[n] += int(r[ix[n]] or 0)
print(agg.most_common(7))
for i, s in enumerate(src):
    if 'SYNCS' in s and 'TRYWAIT' in s:
        print("  wait", s.split('[')[1].split(']')[0], "samples", sum(smp[i:i + 5]), "executed", exe[i])
top = sorted(range(len(R)), key=lambda i: -smp[i])[:14]
for i in top:
    r = R[i]; st = {n: int(r[ix[n]] or 0) for n in stalls}
    print(f"{smp[i]:5d} {src[i][:72]:72s} {sorted(st.items(), key=lambda kv: -kv[1])[:2]}")
