"""Summarise an `ncu --page source --csv` dump: stall samples by opcode and by reason, hottest SASS ranges."""
import csv, collections, re, sys
rows = list(csv.reader(open(sys.argv[1])))
H = rows[1]
si = H.index('Source'); ai = H.index('Warp Stall Sampling (All Samples)'); ei = H.index('Instructions Executed')
def I(v):
    try: return int(v)
    except ValueError: return 0
data = [r for r in rows[2:] if len(r) > max(ai, ei)]
tot = sum(I(r[ai]) for r in data); totexec = sum(I(r[ei]) for r in data)
print('sass instrs', len(data), 'samples', tot, 'warp-instr executed', totexec)
op = collections.Counter(); ex = collections.Counter()
for r in data:
    m = re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)', r[si]); o = m.group(2).split('.')[0] if m else '?'
    op[o] += I(r[ai]); ex[o] += I(r[ei])
for o, c in op.most_common(24): print(f"{o:10s} samples {100*c/tot:5.1f}%  exec {100*ex[o]/totexec:5.1f}%")
for i, h in enumerate(H):
    if h.startswith('stall_') and 'Not Issued' not in h:
        s = sum(I(r[i]) for r in data)
        if s > 0.005 * tot: print(f"{h:26s} {100*s/tot:5.1f}%")
# hottest 64-instruction windows
w = 64
best = []
for a in range(0, len(data), w):
    s = sum(I(r[ai]) for r in data[a:a+w]); e = sum(I(r[ei]) for r in data[a:a+w])
    best.append((s, a, e))
for s, a, e in sorted(best, reverse=True)[:8]:
    print(f"window @{a:5d}: samples {100*s/tot:4.1f}% exec {100*e/totexec:4.1f}%  first: {data[a][si].strip()[:60]}")
