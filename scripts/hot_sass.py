"""Print the hot SASS regions of the first kernel in an ncu report: python scripts/hot_sass.py rep [min_pct]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; minpct = float(sys.argv[2]) if len(sys.argv) > 2 else 0.8
raw = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
out = []; k = 0
for r in rows:
    if r and r[0] == 'Kernel Name':
        k += 1
        if k == 2: break
        continue
    if r and r[0] == 'Address': hdr = r; continue
    if k == 1 and len(r) > 6: out.append(r)
ie = hdr.index('Instructions Executed'); ss = hdr.index('Warp Stall Sampling (All Samples)')
tot = sum(int(r[ie]) for r in out); stot = sum(int(r[ss]) for r in out)
print('total warp instructions', tot, 'SASS lines', len(out), 'stall samples', stot)
prev = -2
for i, r in enumerate(out):
    if int(r[ie]) > tot * minpct / 100 or int(r[ss]) > stot * minpct / 100:
        if i != prev + 1: print('   ...')
        print('%4d %-72s %6.2f%% inst  %6.2f%% stall' % (i, r[1].strip()[:72], 100 * int(r[ie]) / tot, 100 * int(r[ss]) / stot))
        prev = i
