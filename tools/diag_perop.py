"""GPU diagnostic: per-op times of one bs64 641 px forward (CUDA events between the ops), for A/B runs of planner
switches given in the environment (PIFPAF_*)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from openpifpaf_b200 import network     # noqa: E402

b = int(os.environ.get('DIAG_BATCH', '64'))
base = os.environ.get('DIAG_BASE', 'shufflenetv2k16')
plan = network.random_plan(base, seed=7)
x = torch.randn(b, 3, 641, 641, generator=torch.Generator().manual_seed(1)).cuda()
net = network.CompiledNet(plan, 641, 641, b)
for _ in range(3):
    net.forward(x)
torch.cuda.synchronize()
acc = None
R = 8
for _ in range(R):
    ms, kind, flops, nbytes = net.forward_timed(x)
    acc = ms if acc is None else acc + ms
ms = acc / R
tag = ' '.join(f'{k}={v}' for k, v in sorted(os.environ.items()) if k.startswith('PIFPAF_'))
print(f'PEROP [{tag}] b={b}: forward {ms.sum():.3f} ms  gemm {ms[kind == 1].sum():.3f}  dw {ms[kind == 2].sum():.3f}  stem {ms[kind == 0].sum():.3f}', flush=True)
names = {0: 'in', 1: 'g', 2: 'dw', 3: 'fu'}
print('   ' + ' '.join(f'{i}{names[int(kind[i])]}:{ms[i]:.3f}' for i in range(len(ms))), flush=True)
