"""Per-kernel counts of the SASS mnemonics that prove tcgen05 / TMEM / TMA use (B200_PROFILING.md) in libkgc_b200.so.

    python profiles/sass_summary.py > profiles/r02_sass_summary.md
"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, 'kgc-gcn_b200', 'libkgc_b200.so')
PAT = ['UTCHMMA', 'UTCQMMA', 'UTCBAR', 'LDTM', 'STTM', 'UTMALDG', 'UTMASTG', 'UTMAREDG', 'UBLKCP', 'SYNCS', 'RED.E.ADD.F32', 'ATOM.E.ADD.F32',
       'ATOMG.E.ADD.F32', 'LDG.E.128', 'STG.E.128', 'HMMA', 'IMMA']


def main():
    out = subprocess.run(['cuobjdump', '-sass', SO], stdout=subprocess.PIPE, text=True, check=True).stdout
    counts, cur, arch = collections.OrderedDict(), None, set()
    for line in out.splitlines():
        m = re.search(r'Function : (\S+)', line)
        if m:
            name = subprocess.run(['c++filt', m.group(1)], stdout=subprocess.PIPE, text=True).stdout.strip()
            name = re.sub(r'\(anonymous namespace\)::', '', name).split('(')[0].replace('void ', '').replace('kgc::', '')
            cur = counts.setdefault(name, collections.Counter())
            continue
        m = re.search(r'arch = (sm_\w+)', line)
        if m:
            arch.add(m.group(1))
        if cur is None:
            continue
        for p in PAT:
            if re.search(r'\b' + re.escape(p), line):
                cur[p] += 1
    cols = [p for p in PAT if any(c[p] for c in counts.values())]
    print('# SASS evidence: libkgc_b200.so (`cuobjdump -sass`, arch {})\n'.format(', '.join(sorted(arch))))
    print('Counts of instruction occurrences per kernel (static code, not executions).  `UTCHMMA` = tcgen05.mma (kind::f16 / tf32),')
    print('`LDTM` / `STTM` = tcgen05.ld / tcgen05.st (tensor memory), `UTMALDG` / `UTMASTG` = TMA tensor load / store, `UTCBAR` =')
    print('tcgen05.commit, `SYNCS` = mbarrier operations.  No kernel holds a float atomic (`RED/ATOM ... ADD.F32` columns absent')
    print('or zero): every reduction is ordered.\n')
    print('| kernel | ' + ' | '.join(cols) + ' |')
    print('|---|' + '---:|' * len(cols))
    tot = collections.Counter()
    for name, c in counts.items():
        tot.update(c)
        if any(c[p] for p in cols):
            print('| `{}` | '.format(name[:70]) + ' | '.join(str(c[p]) if c[p] else '' for p in cols) + ' |')
    print('| **all {} kernels** | '.format(len(counts)) + ' | '.join('**{}**'.format(tot[p]) for p in cols) + ' |')
    fa = sum(tot[p] for p in ('RED.E.ADD.F32', 'ATOM.E.ADD.F32', 'ATOMG.E.ADD.F32'))
    print('\nFloat atomics in the whole library: **{}**.'.format(fa))


if __name__ == '__main__':
    main()
