"""Aggregate an `ncu --page source --csv --print-source cuda,sass` dump by CUDA source line."""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
cur = None; agg = collections.Counter(); aggx = collections.Counter(); ai = ei = None
for r in rows:
    if not r: continue
    if r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if r[0] == 'Function Name': continue
    if r[0] == 'Line No':
        ai = r.index('Warp Stall Sampling (All Samples)'); ei = r.index('Instructions Executed') if 'Instructions Executed' in r else None; continue
    if r[0].isdigit():
        def I(v):
            try: return int(v)
            except ValueError: return 0
        key = (cur, int(r[0]), r[1].strip()[:96])
        agg[key] += I(r[ai]); aggx[key] += I(r[ei]) if ei is not None else 0
tot = sum(agg.values()) or 1; totx = sum(aggx.values()) or 1
byf = collections.Counter(); byfx = collections.Counter()
for k, v in agg.items(): byf[k[0]] += v; byfx[k[0]] += aggx[k]
print("# stall samples / executed warp-instructions by file")
for f, v in byf.most_common(): print(f"{f:28s} samples {100*v/tot:5.1f}%  exec {100*byfx[f]/totx:5.1f}%")
print("# hottest source lines")
for k, v in agg.most_common(top): print(f"{100*v/tot:5.1f}%  x{100*aggx[k]/totx:5.1f}%  {k[0]}:{k[1]}  {k[2]}")
