#!/usr/bin/env python
"""SASS opcode histogram per object of the CUDA library (what proves the Blackwell-native path: UTC*MMA = tcgen05.mma,
LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG/UBLKCP = TMA, HMMA = legacy mma.sync, MUFU.* = SFU):
    python tools/sass_summary.py [profiles/sass_summary.txt]
Runs here (no GPU): cuobjdump -sass rep-yolo_b200/csrc/_obj/*.o, built with -gencode arch=compute_100a,code=sm_100a."""
import collections
import glob
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ['UTCHMMA', 'UTCQMMA', 'UTCBAR', 'UTCCP', 'UTCATOMSWS', 'LDTM', 'STTM', 'UTMALDG', 'UTMASTG', 'UBLKCP', 'UTMAPF', 'UTMACCTL', 'UTMACMDFLUSH',
        'SYNCS', 'HMMA', 'LDSM', 'LDGSTS', 'MUFU.TANH', 'MUFU.EX2', 'MUFU.RCP', 'REDUX', 'MATCH', 'VOTE', 'ELECT', 'ACQBULK', 'NANOSLEEP', 'BAR.SYNC',
        'ERRBAR', 'CCTL']


def main():
    out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, 'profiles', 'sass_summary.txt')
    lines = ['# cuobjdump -sass opcode counts per object (sm_100a); kernels listed with their instruction totals',
             '# tcgen05.mma -> UTCHMMA, tcgen05.ld -> LDTM, tcgen05.commit -> UTCBAR, TMA tensor load/store -> UTMALDG/UTMASTG, '
             'cp.async.bulk -> UBLKCP, mbarrier -> SYNCS, mma.sync -> HMMA, ldmatrix -> LDSM, cp.async -> LDGSTS']
    for obj in sorted(glob.glob(os.path.join(ROOT, 'rep-yolo_b200', 'csrc', '_obj', '*.o'))):
        txt = subprocess.run(['cuobjdump', '-sass', obj], capture_output=True, text=True).stdout
        per_kernel, cur = collections.OrderedDict(), None
        for ln in txt.splitlines():
            m = re.search(r'Function : (\S+)', ln)
            if m:
                cur = per_kernel.setdefault(m.group(1), collections.Counter())
                continue
            m = re.match(r'\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', ln)
            if m and cur is not None:
                op = m.group(1)
                cur['_total'] += 1
                for k in KEYS:
                    if op == k or op.startswith(k + '.') or (k.count('.') and op.startswith(k)):
                        cur[k] += 1
        lines.append(f'\n== {os.path.basename(obj)}')
        tot = collections.Counter()
        for c in per_kernel.values():
            tot.update(c)
        lines.append('  all kernels: ' + ', '.join(f'{k} {tot[k]}' for k in KEYS if tot[k]) + f' (instructions {tot["_total"]})')
        for name, c in per_kernel.items():
            short = re.sub(r'^_ZN2ry\d+_GLOBAL__N__\w+?_cu_[0-9a-f]+', '', name)[:70]
            hot = ', '.join(f'{k} {c[k]}' for k in KEYS if c[k])
            lines.append(f'  {short:70s} {c["_total"]:6d} instr: {hot}')
    open(out, 'w').write('\n'.join(lines) + '\n')
    print('\n'.join(lines[:40]))


if __name__ == '__main__':
    main()
