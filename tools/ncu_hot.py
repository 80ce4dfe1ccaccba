"""Hot SASS lines of one launch of an .ncu-rep (source page): samples, executed count and top stall reasons per instruction.
usage: python tools/ncu_hot.py REP LAUNCH_INDEX [N]"""
import collections, csv, subprocess, sys
rep, idx = sys.argv[1], int(sys.argv[2])
n = int(sys.argv[3]) if len(sys.argv) > 3 else 40
raw = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--launch-skip', str(idx), '--launch-count', '1'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
print(rows[0][1][:100])
h = rows[1]
ia, isrc, ismp, iex = h.index('Address'), h.index('Source'), h.index('# Samples'), h.index('Instructions Executed')
stall = [(j, k.replace('stall_', '')) for j, k in enumerate(h) if k.startswith('stall_') and 'Not Issued' not in k]
sass = [r for r in rows[2:] if len(r) > iex and r[ia].startswith('0x')]
# the csv has the SASS view first, then (after another header) the source-line view
end = [j for j, r in enumerate(rows) if j > 2 and r and r[0] == 'Address']
if end:
    sass = [r for r in rows[2:end[0]] if len(r) > iex and r[ia].startswith('0x')]
tot = sum(int(r[ismp]) for r in sass) or 1
print('total samples', tot, 'instructions executed', sum(int(r[iex]) for r in sass))
agg = collections.Counter()
for r in sass:
    for j, k in stall:
        agg[k] += int(r[j] or 0)
print('stalls:', ', '.join('%s %.1f%%' % (k, 100.0 * v / tot) for k, v in agg.most_common(10)))
op = collections.Counter(); opx = collections.Counter()
for r in sass:
    t = r[isrc].split()
    o = (t[1] if t[0].startswith('@') else t[0]).split('.')[0]
    op[o] += int(r[ismp]); opx[o] += int(r[iex])
print('opcodes by samples:', ', '.join('%s %.1f%%' % (k, 100.0 * v / tot) for k, v in op.most_common(12)))
print('opcodes by executed:', ', '.join('%s %d' % (k, v) for k, v in opx.most_common(14)))
top = sorted(range(len(sass)), key=lambda i: -int(sass[i][ismp]))[:n]
for i in sorted(top):
    r = sass[i]
    st = sorted(((int(r[j] or 0), k) for j, k in stall), reverse=True)[:3]
    print('%5d %-10s smp %6d (%.1f%%) exe %9s  %-70s %s' % (i, r[ia][-6:], int(r[ismp]), 100.0 * int(r[ismp]) / tot, r[iex], r[isrc][:70], ' '.join('%s:%d' % (k, v) for v, k in st if v)))
