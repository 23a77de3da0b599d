"""Static vs dynamic instruction footprint of every kernel in an ncu report (source page, SASS view).
usage: sass_footprint.py report.ncu-rep"""
import csv, collections, io, subprocess, sys
raw = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
secs = []; cur = None; hdr = None
for r in csv.reader(io.StringIO(raw)):
    if not r: continue
    if r[0] == 'Kernel Name': cur = {'name': r[1], 'rows': []}; secs.append(cur); continue
    if r[0] == 'Address': hdr = r; continue
    if cur is not None and hdr and len(r) == len(hdr): cur['rows'].append(r)
ei = hdr.index('Instructions Executed'); si = hdr.index('Source'); wi = hdr.index('Warp Stall Sampling (All Samples)')
for s in secs:
    data = s['rows']; n = len(data); ex = [int(r[ei]) for r in data]; tot = sum(ex); warps = max(ex[:8])
    print(f"\n== {s['name'][:100]}\nstatic {n} instrs = {n*16/1024:.1f} KB; executed {tot} warp-instrs = {tot/warps:.0f} per warp of the first block")
    print(f"instructions executed by >= half of the warps: {sum(e >= 0.5*warps for e in ex)} ({sum(e >= 0.5*warps for e in ex)*16/1024:.1f} KB); never executed: {sum(e == 0 for e in ex)}")
    ops = collections.Counter(); opsd = collections.Counter()
    for r, e in zip(data, ex):
        t = r[si].split(); op = (t[1] if t[0].startswith('@') else t[0]).split('.')[0]; ops[op] += 1; opsd[op] += e
    print("  ".join(f"{op}:{100*c/tot:.1f}%" for op, c in opsd.most_common(12)))
    print("cumulative share of executed instructions / stall samples per 512-instruction (8 KB) chunk:")
    print("  ".join(f"[{i//512}] {100*sum(ex[i:i+512])/tot:.0f}%/{sum(int(r[wi]) for r in data[i:i+512])}" for i in range(0, n, 512)))
