"""Print SASS lines of an `ncu --page source --csv` dump whose executed count is within
[lo, hi] x the given per-warp-step normaliser (to look at the hot loop)."""
import csv, sys
path, per_step = sys.argv[1], float(sys.argv[2])
lo = float(sys.argv[3]) if len(sys.argv) > 3 else 0.05
rows = list(csv.reader(open(path)))
hdr = rows[1]
iS, iE, iT, iA = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("Address")
iSamp = hdr.index("# Samples")
for r in rows[2:]:
    if len(r) <= iT or not r[iE]: continue
    n = int(r[iE])
    if n / per_step >= lo:
        print(f"{r[iA][-5:]} {n/per_step:6.3f} thr={int(r[iT])/max(n,1):4.1f} smp={r[iSamp]:>6s}  {r[iS].strip()}")
