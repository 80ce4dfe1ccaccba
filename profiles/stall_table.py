"""Per-kernel warp-stall breakdown from the source page of an .ncu-rep (19 launches of one forward, in order).
usage: python profiles/stall_table.py gpurun_out/prof_XXX.ncu-rep profiles/NAME_stalls.md"""
import collections
import csv
import subprocess
import sys

NAMES = ['conv%d' % i for i in range(14)] + ['decode_conv%d' % j for j in range(1, 6)]                          # round 1: 19 launches
NAMES_R2 = ['conv%d' % i for i in range(7)] + ['conv7..conv11', 'conv12', 'conv13'] + ['decode_conv%d' % j for j in range(1, 6)]   # chain kernel: 15


def one(rep, i):
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--launch-skip', str(i), '--launch-count', '1'],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    h = rows[1]
    ia, isrc, ismp, iex = h.index('Address'), h.index('Source'), h.index('# Samples'), h.index('Instructions Executed')
    stall = [(j, k) for j, k in enumerate(h) if k.startswith('stall_') and 'Not Issued' not in k]
    end = [j for j, r in enumerate(rows) if j > 2 and r and r[0] == 'Address']
    sass = [r for r in rows[2:(end[0] if end else len(rows))] if len(r) > iex and r[ia].startswith('0x')]
    tot = sum(int(r[ismp]) for r in sass) or 1
    st, op = collections.Counter(), collections.Counter()
    for r in sass:
        n = int(r[ismp])
        if not n:
            continue
        t = r[isrc].split()
        op[t[1] if t[0].startswith('@') else t[0]] += n
        for j, k in stall:
            st[k.replace('stall_', '')] += int(r[j] or 0)
    instr = sum(int(r[iex]) for r in sass)
    return tot, instr, st, op


def main(rep, out):
    with open(out, 'w') as f:
        f.write('# Warp-stall sampling per kernel (ncu source page, one forward, fp16 b64 224x224)\n\n'
                'Shares of all warp samples of the kernel (every warp of every role, so waiting roles count: `long_sb` and\n'
                '`barrier` are mostly warps parked on an mbarrier or the final `__syncthreads`, `selected` + `not_selected` are\n'
                'warps that could issue).  Opcode column: where the samples sit.\n\n'
                '| stage | samples | warp-instr | top stall reasons (%) | top opcodes by samples (%) |\n|---|---:|---:|---|---|\n')
        for i, name in enumerate(NAMES_R2 if len(sys.argv) > 3 and sys.argv[3] == 'r2' else NAMES):
            tot, instr, st, op = one(rep, i)
            f.write('| %s | %d | %.1f M | %s | %s |\n' % (
                name, tot, instr / 1e6,
                ', '.join('%s %d' % (k, round(100 * v / tot)) for k, v in st.most_common(5)),
                ', '.join('%s %d' % (k, round(100 * v / tot)) for k, v in op.most_common(4))))
    print('wrote', out)


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2])
