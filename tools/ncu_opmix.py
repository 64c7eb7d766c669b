"""Executed warp instructions of one kernel launch by opcode and by issue pipe, and the source lines that issue the most
ALU-pipe instructions -- from `ncu -i report.ncu-rep --page source --csv --print-source cuda,sass --launch-skip N
--launch-count 1` (no GPU needed; the report must come from the same source revision for the line texts to fit).

    python tools/ncu_opmix.py src.csv [lines]

The ALU pipe (LOP3, ISETP, SHF, SEL, IADD3, PRMT, PLOP3, LEA ...) runs at half rate; IMAD and friends go to the FMA
pipes.  A kernel made of integer tests is bound by ALU-pipe instructions, not by the issue rate."""
import csv
import re
import sys
from collections import Counter, defaultdict

ALU = {'LOP3', 'ISETP', 'SHF', 'SEL', 'VIADD', 'IADD3', 'PRMT', 'PLOP3', 'LEA', 'IABS', 'VIMNMX', 'FLO', 'BREV', 'MOV', 'CS2R',
       'SGXT', 'BMSK', 'P2R', 'R2P', 'FSEL', 'FSETP', 'IMNMX', 'VIADDMNMX'}
LSU = {'LDG', 'STG', 'LDS', 'STS', 'ATOMS', 'ATOMG', 'RED', 'REDG', 'LDL', 'STL', 'LDC', 'LDCU', 'ATOM', 'CCTL'}
CTRL = {'BRA', 'BSSY', 'BSYNC', 'EXIT', 'WARPSYNC', 'NOP', 'CALL', 'RET', 'BREAK', 'BAR', 'PREEXIT', 'ACQBULK', 'YIELD', 'BRX'}
WARP = {'SHFL', 'VOTE', 'REDUX', 'MATCH', 'S2R', 'S2UR', 'R2UR', 'VOTEU'}


def pipe(op):
    b = op.split('.')[0]
    if b in ALU:
        return 'alu'
    if b == 'IMAD':
        return 'fma-heavy' if ('WIDE' in op or 'HI' in op or op in ('IMAD', 'IMAD.X')) else 'fma'
    if b in ('FFMA', 'FMUL', 'FADD'):
        return 'fma'
    if b in LSU:
        return 'lsu'
    if b in CTRL:
        return 'control'
    if b in WARP:
        return 'warp / special'
    return 'other'


rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hi = [i for i, r in enumerate(rows) if r and r[0] == 'Line No']
h = rows[hi[0]]
ci = h.index('Instructions Executed')
end = hi[1] - 2 if len(hi) > 1 else len(rows)
by_op, by_pipe, by_line, text, seen, cur = Counter(), Counter(), defaultdict(Counter), {}, set(), None
for r in rows[hi[0] + 1:end]:
    if len(r) <= ci:
        continue
    if r[0] != '':
        try:
            cur = int(r[0]); text[cur] = r[1]
        except ValueError:
            cur = None
        continue
    if r[2] in seen:                   # an inlined instruction is listed under every enclosing source line
        continue
    seen.add(r[2])
    m = re.match(r'(@!?U?P\d+\s+)?([A-Z0-9_.]+)', r[3].strip())
    if not m:
        continue
    try:
        n = float(r[ci] or 0)
    except ValueError:
        n = 0.0
    op = m.group(2)
    by_op[op.split('.')[0]] += n
    by_pipe[pipe(op)] += n
    by_line[cur][pipe(op)] += n
tot = sum(by_op.values())
print(rows[1][1][:110] if len(rows) > 1 and len(rows[1]) > 1 else '')
print(f'executed warp instructions: {tot / 1e6:.1f} M')
for k, v in by_pipe.most_common():
    print(f'  {k:16s} {v / 1e6:8.2f} M {100 * v / tot:5.1f} %')
print('opcodes: ' + ', '.join(f'{k} {100 * v / tot:.1f} %' for k, v in by_op.most_common(14)))
print('source lines by ALU-pipe instructions:')
for ln, c in sorted(by_line.items(), key=lambda kv: -kv[1]['alu'])[:top]:
    print(f'  {ln:5d}  alu {c["alu"] / 1e6:6.2f} M  all {sum(c.values()) / 1e6:6.2f} M | {text.get(ln, "").strip()[:100]}')
