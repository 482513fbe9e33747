"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: one MGCNConv fwd+bwd step (eager launches)."""
import collections
import csv
import re
import sys


def main(path, which=5):
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
    hdr, data = rows[hi], rows[hi + 1:]
    kn, mv, idc = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('ID')

    def short(n):
        n = n.replace('kgc::<unnamed>::', '').replace('void ', '')
        return re.sub(r'\(.*', '', n)[:84]
    seq = [(int(r[idc]), short(r[kn]), float(r[mv].replace(',', ''))) for r in data]
    idx = [i for i, s in enumerate(seq) if s[1].startswith(('agg_lean_kernel<0', 'agg_stream_kernel<0', 'agg_fwd_kernel'))]
    # a step = the launches between two consecutive forward aggregation kernels (shifted to the step's first launch)
    first = next(i for i, s in enumerate(seq) if s[1].startswith(('pack_b_kernel', 'conv_prep_kernel')) or 'distribution_elementwise' in s[1])
    shift = idx[0] - first if first < idx[0] else 0
    a, b = idx[which] - shift, idx[which + 1] - shift
    step = seq[a:b]
    tot = sum(s[2] for s in step)
    agg = collections.OrderedDict()
    for _, n, t in step:
        agg.setdefault(n, [0, 0.0])
        agg[n][0] += 1
        agg[n][1] += t
    print('| kernel | launches | total us | share |\n|---|---:|---:|---:|')
    for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print('| `{}` | {} | {:.1f} | {:.1%} |'.format(n, c, t / 1e3, t / tot))
    print('| **total** | {} | {:.1f} | 100% |'.format(len(step), tot / 1e3))


if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 5)
