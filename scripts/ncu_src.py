"""Summarise an `ncu --page source --csv` dump: opcode mix and top stall locations."""
import csv, collections, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
data = [r for r in rows[2:] if len(r) == len(hdr) and r[0].startswith('0x')]
if '--first' in sys.argv:
    # several kernels in one dump: keep the first
    cut = next((k for k, r in enumerate(rows[2:]) if r and r[0] == 'Kernel Name'), None)
    if cut is not None:
        data = [r for r in rows[2:2 + cut] if len(r) == len(hdr) and r[0].startswith('0x')]
iS = hdr.index('Source'); iN = hdr.index('# Samples'); iE = hdr.index('Instructions Executed')
tot = sum(int(r[iN]) for r in data); totE = sum(int(r[iE]) for r in data)
print("instr rows", len(data), "samples", tot, "warp instrs", totE)
op = collections.Counter(); ops = collections.Counter()
for r in data:
    m = re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)', r[iS]); o = m.group(2).split('.')[0] if m else '?'
    op[o] += int(r[iE]); ops[o] += int(r[iN])
for o, c in op.most_common(28):
    print("%-10s exec=%5.1f%% samples=%5.1f%%" % (o, 100 * c / totE, 100 * ops[o] / tot))
print()
stall_cols = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
agg = collections.Counter()
for r in data:
    for i in stall_cols: agg[hdr[i]] += int(r[i])
print("stall totals:", [(k, "%.1f%%" % (100 * v / tot)) for k, v in agg.most_common(8)])
for r in sorted(data, key=lambda r: -int(r[iN]))[:int(sys.argv[2]) if len(sys.argv) > 2 else 22]:
    st = sorted(((int(r[i]), hdr[i]) for i in stall_cols), reverse=True)[:2]
    print(r[0][-5:], r[iN], r[iE], r[iS].strip()[:64], st)
