"""ncu target: `python tools/prof_target.py net|dec [fuse]` -- two bs64 forwards (641 px, shufflenetv2k16) or two
planted bs64 decodes + one crowd bs8 decode.  Not product code."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from openpifpaf_b200 import decoder, network, synth      # noqa: E402

what = sys.argv[1]
if what == 'net':
    fuse = len(sys.argv) > 2 and sys.argv[2] == '1'
    plan = network.random_plan('shufflenetv2k16', seed=0)
    net = network.CompiledNet(plan, 641, 641, 64, fuse_dw=fuse)
    x = torch.randn(64, 3, 641, 641, generator=torch.Generator().manual_seed(1)).cuda()
    for _ in range(2):
        net.forward(x)
    torch.cuda.synchronize()
    print('ops per forward:', len(net.op_desc), [o['kind'] for o in net.op_desc])
else:
    for workload, B, people in (('cocokp', 64, None), ('cocokp', 8, 30)):
        batch = synth.make_batch(workload, B, 41, 41, people, seed=11)
        d = decoder.CifCaf(batch['n_keypoints'], torch.from_numpy(batch['skeleton']))
        cif = torch.from_numpy(batch['cif']).cuda()
        caf = torch.from_numpy(batch['caf']).cuda()
        for _ in range(2):
            res = d.decode_batch(cif, 16, caf, 16)
        print(workload, B, people, sum(len(a) for a, _ in res), 'annotations')
