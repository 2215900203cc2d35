"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel (ms, share)."""
import collections
import csv
import re
import sys


def summarize(path):
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
    idx = {h: i for i, h in enumerate(rows[hi])}
    agg = collections.OrderedDict()
    for r in rows[hi + 2:]:
        if len(r) < len(idx) or r[idx['Metric Name']] != 'gpu__time_duration.sum':
            continue
        name = re.sub(r'<unnamed>::|\(.*|void ', '', r[idx['Kernel Name']])
        v = float(r[idx['Metric Value']].replace(',', ''))
        unit = r[idx['Metric Unit']]
        v = v / 1e6 if unit == 'ns' else (v / 1e3 if unit.startswith('us') else v)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(v[1] for v in agg.values())
    lines = [f'total {tot:.3f} ms over {sum(v[0] for v in agg.values())} launches']
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append(f'{k[:60]:60s} {n:5d} launches {t:10.3f} ms  {100 * t / tot:5.1f}%  avg {t / n * 1e3:9.1f} us')
    return '\n'.join(lines)


if __name__ == '__main__':
    for p in sys.argv[1:]:
        print('==', p)
        print(summarize(p))
