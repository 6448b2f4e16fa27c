"""Executed-instruction mix by opcode for the first kernel of an ncu report: python scripts/sass_mix.py rep"""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]
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
mix = collections.Counter(); st = collections.Counter()
for r in out:
    t = r[1].strip().split()
    op = t[1] if t[0].startswith('@') else t[0]
    op = op.split('.')[0]
    mix[op] += int(r[ie]); st[op] += int(r[ss])
tot = sum(mix.values()); stot = sum(st.values())
print('total', tot)
for op, n in mix.most_common(28):
    print('%-12s %6.2f%% inst %6.2f%% stall' % (op, 100 * n / tot, 100 * st[op] / stot))
