"""Summarise an `ncu --page source --csv` dump: executed warp instructions by opcode class."""
import csv, sys, collections
path = sys.argv[1]
per_step = float(sys.argv[2]) if len(sys.argv) > 2 else None  # warp-steps in the launch
rows = list(csv.reader(open(path)))
hdr = rows[1]
iS, iE, iT = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
iSamp = hdr.index("# Samples")
ops = collections.Counter(); thr = collections.Counter(); samp = collections.Counter()
tot = 0
for r in rows[2:]:
    if len(r) <= iT or not r[iE]: continue
    n = int(r[iE]); t = int(r[iT])
    src = r[iS].strip()
    toks = src.split()
    op = toks[1] if toks and toks[0].startswith('@') else (toks[0] if toks else '?')
    base = op.split('.')[0]
    cls = op if base.startswith('MUFU') else base
    ops[cls] += n; thr[cls] += t; samp[cls] += int(r[iSamp] or 0)
    tot += n
print(f"total warp instructions {tot:.3e}" + (f"  = {tot/per_step:.1f} per warp-step" if per_step else ""))
for k, v in ops.most_common(45):
    print(f"{k:18s} {v:14d} {100*v/tot:6.2f}%  avg_thr {thr[k]/max(v,1):5.1f}" + (f"  per-step {v/per_step:7.2f}" if per_step else "") + f"  samples {samp[k]}")
