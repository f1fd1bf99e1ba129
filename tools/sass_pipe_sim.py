"""Pipe-occupancy replay of a kernel's SASS (no GPU needed): classifies every instruction of the longest straight-line block
(the unrolled row body of comb_ws_kernel) by issue pipe and replays it on a scheduler model in which the ALU pipe
(HSET2/PRMT/LOP3/IADD3/VABSDIFF4/ISETP/SEL...) and the FMA-heavy pipe (HFMA2/HADD2/IMAD/IDP/VIADD) each accept one warp
instruction every two cycles (measured with tools/pipe_probe.cu) and one instruction issues per cycle.  Prints the
instruction mix, the pipe-bound ceiling and the issue rate of 1..4 phase-shifted warps per scheduler.

    python tools/sass_pipe_sim.py <file.cubin|lib.so> [mangled-kernel-substring] [--show]
"""
import random
import re
import subprocess
import sys

F_OPS = ('HFMA2', 'HADD2', 'HMUL2', 'IMAD', 'IDP', 'FFMA', 'FADD', 'FMUL', 'VIADD')
A_OPS = ('HSET2', 'PRMT', 'LOP3', 'IADD3', 'VABSDIFF4', 'SEL', 'ISETP', 'MOV', 'PLOP3', 'SHF', 'LEA', 'IABS', 'VIMNMX', 'VIMNMX3',
         'HSETP2', 'IADD', 'HMNMX2', 'FSET', 'FSETP', 'FMNMX', 'I2FP', 'F2FP')
L_OPS = ('LDS', 'LDSM', 'STS', 'LDG', 'STG', 'RED', 'ATOM', 'ATOMG', 'REDUX', 'LDL', 'STL', 'LDC', 'LDCU', 'LDTM')


def pipe(op):
    b = op.split('.')[0]
    return 'F' if b in F_OPS else 'A' if b in A_OPS else 'L' if b in L_OPS else 'O'


def sim(nw, seq, cycles=120000, seed=0):
    random.seed(seed)
    pc = [random.randrange(len(seq)) for _ in range(nw)]
    free = {'A': 0, 'F': 0, 'O': 0, 'L': 0}
    issued, last = 0, 0
    for t in range(cycles):
        for k in range(nw):
            w = (last + 1 + k) % nw
            p = seq[pc[w]]
            if free[p] <= t:
                free[p] = t + (2 if p in 'AF' else 1)
                pc[w] = (pc[w] + 1) % len(seq)
                issued += 1
                last = w
                break
    return issued / cycles


def load(path, kernel=None):
    cmd = ['cuobjdump', '-sass', path]
    out = subprocess.run(cmd, capture_output=True, text=True).stdout
    ops, take = [], kernel is None
    for line in out.splitlines():
        if 'Function :' in line:
            take = kernel is None or kernel in line
        m = re.match(r'\s+/\*[0-9a-f]{4,5}\*/\s+(.*?);', line)
        if m and take:
            t = m.group(1).split()
            ops.append(t[1] if t[0].startswith('@') else t[0])
    return ops


if __name__ == '__main__':
    args = [a for a in sys.argv[1:] if not a.startswith('--')]
    ops = load(args[0], args[1] if len(args) > 1 else None)
    s = ''.join(pipe(o) for o in ops)
    cuts = [-1] + [i for i, o in enumerate(ops) if o.split('.')[0] in ('BRA', 'BSSY', 'BSYNC', 'EXIT', 'CALL', 'RET', 'WARPSYNC', 'SYNCS', 'ENDCOLLECTIVE')] + [len(ops)]
    lo, hi = max(((cuts[i] + 1, cuts[i + 1]) for i in range(len(cuts) - 1)), key=lambda p: p[1] - p[0])
    seq = list(s[lo:hi])
    cnt = {k: seq.count(k) for k in 'AFLO'}
    print('instructions', len(ops), '| longest block', len(seq), cnt, '| pipe-bound ceiling %.3f' % (len(seq) / max(2 * cnt['A'], 2 * cnt['F'], len(seq))))
    for nw in (1, 2, 3, 4):
        print(nw, 'warps per scheduler: issue rate', ' '.join('%.3f' % sim(nw, seq, seed=x) for x in range(3)))
    if '--show' in sys.argv:
        for i in range(lo, hi, 120):
            print(s[i:i + 120])
