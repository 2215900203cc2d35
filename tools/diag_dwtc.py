"""GPU diagnostic: depthwise 5x5 on the tensor cores (k_dwconv5_tc) against the FMA kernel (k_dwconv5_tma).

Part 1 (small net, every depthwise output tapped): which UMMA descriptor variant reproduces the FMA kernel --
window pitch 12 / 16 pixels, base-offset field 0 / (start >> 7) & 7.  The tensor-core path rounds the depthwise
weights to bf16, so "equal" means a relative error of a few 2^-9, a wrong descriptor gives O(1).
Part 2 (641 px, batch 64): per-op times of both kernels.

Run under a short timeout first: a protocol bug in a warp-specialised kernel shows up as a hang."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from openpifpaf_b200 import network     # noqa: E402

MODE = sys.argv[1] if len(sys.argv) > 1 else 'all'
TC = int(os.environ.get('DIAG_DW_TC', '1'))     # bit 0 stride 1, bit 1 stride 2


def build(plan, h, w, b, tc, pwid=12, bo=0):
    os.environ['PIFPAF_DW_TC'] = str(tc)
    os.environ['PIFPAF_DW_TC_PWID'] = str(pwid)
    os.environ['PIFPAF_DW_TC_BO'] = str(bo)
    return network.CompiledNet(plan, h, w, b, fuse_dw=False)


def dw_outputs(net, b):
    out = {}
    for o in net.op_desc:
        if o['kind'] == 'dwconv':
            c = o['channels']
            out[(o['out'], o['stride'])] = net.tap(o['out'], b)[..., o['out_off']:o['out_off'] + c]
    return out


if MODE in ('all', 'small'):
    plan = network.random_plan('shufflenetv2k16', seed=7)
    for (h, w, b) in ((97, 129, 2), (337, 401, 3)):
        x = torch.randn(b, 3, h, w, generator=torch.Generator().manual_seed(12)).cuda()
        ref_net = build(plan, h, w, b, 0)
        ref_heads = [t.clone() for t in ref_net.forward(x)]
        torch.cuda.synchronize()
        ref = dw_outputs(ref_net, b)
        for pwid, bo in ((12, 0), (16, 0), (16, 1), (12, 1)):
            net = build(plan, h, w, b, TC, pwid, bo)
            heads = [t.clone() for t in net.forward(x)]
            torch.cuda.synchronize()
            got = dw_outputs(net, b)
            worst = {}
            for k, r in ref.items():
                # only the first depthwise op of each stride sees identical inputs in both nets; later ones differ
                # by the propagated rounding, so report all of them against their own scale
                e = float(np.abs(got[k] - r).max()) / max(float(np.abs(r).max()), 1e-6)
                worst[k[1]] = max(worst.get(k[1], 0.0), e)
            he = max(float((a - b_).abs().max()) for a, b_ in zip(heads, ref_heads))
            print(f'DWTC {h}x{w} b{b} pwid={pwid} bo={bo}: max rel err by stride {worst}  heads max abs {he:.4g}', flush=True)
            net.close()
        ref_net.close()

if MODE in ('all', 'perf'):
    plan = network.random_plan('shufflenetv2k16', seed=7)
    b = 64
    x = torch.randn(b, 3, 641, 641, generator=torch.Generator().manual_seed(1)).cuda()
    pw = int(os.environ.get('DIAG_PWID', '12'))
    bo = int(os.environ.get('DIAG_BO', '0'))
    for tc in (0, TC):
        net = build(plan, 641, 641, b, tc, pw, bo)
        for _ in range(3):
            net.forward(x)
        torch.cuda.synchronize()
        acc = None
        for _ in range(5):
            ms, kind, flops, nbytes = net.forward_timed(x)
            acc = ms if acc is None else acc + ms
        ms = acc / 5
        dw = [(i, float(ms[i]), float(nbytes[i])) for i in range(len(ms)) if kind[i] == 2]
        print(f'PERF tc={tc}: forward {ms.sum():.3f} ms, dw total {sum(m for _, m, _ in dw):.3f} ms, gemm {ms[kind == 1].sum():.3f} ms', flush=True)
        print('   dw ops (ms, GB/s alg): ' + ' '.join(f'{m:.3f}/{nb * b / m / 1e6:.0f}' for _, m, nb in dw), flush=True)
        net.close()
