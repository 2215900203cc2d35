"""Summarise an `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --csv` capture of
ONE bench step: DRAM bytes per kernel (sum over its launches and average per launch) -> JSON for bench.py's
`roofline.traffic`.

usage: python tools/summarize_traffic.py <capture.csv> <out.json> [launches_per_step_of_k_gemm_tc]
"""
import collections
import csv
import json
import re
import sys

UNIT = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'ns': 1e-6, 'us': 1e-3, 'ms': 1.0}


def summarize(path):
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
    idx = {h: i for i, h in enumerate(rows[hi])}
    per_launch = collections.OrderedDict()          # launch id -> {name, read, write, ms}
    for r in rows[hi + 1:]:
        if len(r) < len(idx):
            continue
        name = re.sub(r'<unnamed>::|\(.*|void ', '', r[idx['Kernel Name']])
        rec = per_launch.setdefault(r[idx['ID']], {'name': name, 'read': 0.0, 'write': 0.0, 'ms': 0.0})
        v = float(r[idx['Metric Value']].replace(',', '')) * UNIT.get(r[idx['Metric Unit']], 1.0)
        m = r[idx['Metric Name']]
        if m == 'dram__bytes_read.sum':
            rec['read'] = v
        elif m == 'dram__bytes_write.sum':
            rec['write'] = v
        elif m == 'gpu__time_duration.sum':
            rec['ms'] = v
    return list(per_launch.values())


def main():
    launches = summarize(sys.argv[1])
    out = {'source': sys.argv[1], 'kernels': {}}
    agg = collections.OrderedDict()
    for rec in launches:
        a = agg.setdefault(rec['name'], {'launches': 0, 'dram_read_bytes': 0.0, 'dram_write_bytes': 0.0, 'ms_under_ncu': 0.0})
        a['launches'] += 1
        a['dram_read_bytes'] += rec['read']
        a['dram_write_bytes'] += rec['write']
        a['ms_under_ncu'] += rec['ms']
    for k, a in agg.items():
        a['dram_bytes_per_launch'] = (a['dram_read_bytes'] + a['dram_write_bytes']) / a['launches']
        out['kernels'][k] = a
    json.dump(out, open(sys.argv[2], 'w'), indent=1)
    for k, a in agg.items():
        print('%-40s %4d launches  read %8.3f GB  write %8.3f GB  per launch %8.2f MB' % (
            k[:40], a['launches'], a['dram_read_bytes'] / 1e9, a['dram_write_bytes'] / 1e9, a['dram_bytes_per_launch'] / 1e6))


if __name__ == '__main__':
    main()
