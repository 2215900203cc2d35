"""Algorithmic HBM bytes per image of the lowered op list (no GPU needed): what each activation layout has to
move if every operand is read once and every output written once.  usage: python tools/model_bytes.py [size]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from openpifpaf_b200 import network      # noqa: E402


def op_bytes(tensors, o):
    if o['kind'] == 'input_conv':
        h, w, c = tensors[o['out']]
        return o['in_h'] * o['in_w'] * 3 * 4 + h * w * c * 2
    if o['kind'] == 'conv1x1':
        h, w, _ = tensors[o['in']]
        rows = h * w
        out = o['n_out'] * (3 if o['shuffle_src'] >= 0 else 1)       # shuffle: pass-through half read + re-written
        return rows * (o['k_cols'] + out) * 2
    if o['kind'] == 'dwconv':
        hi, wi, _ = tensors[o['in']]
        ho, wo, _ = tensors[o['out']]
        return (hi * wi + ho * wo) * o['channels'] * 2
    if o['kind'] == 'heads':
        h, w, _ = tensors[o['in']]
        return h * w * (o['k_cols'] * 2 + o['w'].shape[0] * 4)
    raise ValueError(o['kind'])


def main():
    size = int(sys.argv[1]) if len(sys.argv) > 1 else 641
    plan = network.random_plan('shufflenetv2k16', seed=0)
    for layout in ('shuffle', 'bins'):
        tensors, ops, _ = network.build_ops(plan, size, size, layout=layout)
        by_kind = {}
        for o in ops:
            by_kind[o['kind']] = by_kind.get(o['kind'], 0) + op_bytes(tensors, o)
        total = sum(by_kind.values())
        act = sum(h * w * c * 2 for (h, w, c) in tensors)
        print(f'{layout:8s} {len(ops)} ops, {len(tensors)} tensors ({act / 1e6:.0f} MB of activations per image): '
              f'{total / 1e6:.0f} MB per image = ' + ', '.join(f'{k} {v / 1e6:.0f}' for k, v in by_kind.items()))


if __name__ == '__main__':
    main()
