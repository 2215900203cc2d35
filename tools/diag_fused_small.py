"""GPU diagnostic: the fused depthwise -> GEMM kernel on ONE small net (run under a short timeout first: a protocol
bug in a warp-specialised kernel shows up as a hang)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from openpifpaf_b200 import network     # noqa: E402

plan = network.random_plan('shufflenetv2k16', seed=7)
x = torch.randn(1, 3, 97, 129, generator=torch.Generator().manual_seed(12)).cuda()
fused = network.CompiledNet(plan, 97, 129, 1, fuse_dw=True)
plain = network.CompiledNet(plan, 97, 129, 1, fuse_dw=False)
hp = [t.clone() for t in plain.forward(x)]
hf = [t.clone() for t in fused.forward(x)]
torch.cuda.synchronize()
print('FUSED_SMALL equal =', all(torch.equal(a, b) for a, b in zip(hf, hp)), flush=True)
