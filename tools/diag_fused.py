"""GPU diagnostic of the fused depthwise -> GEMM kernel (k_dw_gemm): bitwise equality with the two-kernel schedule on
growing sizes, then per-op timing of both schedules at the bench size."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from openpifpaf_b200 import network

def check(base, B, H, W, seed=7):
    plan = network.random_plan(base, seed=seed)
    x = torch.randn(B, 3, H, W, generator=torch.Generator().manual_seed(12)).cuda()
    fused = network.CompiledNet(plan, H, W, B, fuse_dw=True)
    plain = network.CompiledNet(plan, H, W, B, fuse_dw=False)
    hp = [t.clone() for t in plain.forward(x)]
    torch.cuda.synchronize()
    t0 = time.time()
    hf = [t.clone() for t in fused.forward(x)]
    torch.cuda.synchronize()
    ok = all(torch.equal(a, b) for a, b in zip(hf, hp))
    worst = max(float((a - b).abs().max()) for a, b in zip(hf, hp))
    nan = any(bool(torch.isnan(a).any()) for a in hf)
    print(f'DIAG_FUSED {base} B{B} {H}x{W}: equal={ok} max|d|={worst:.4g} nan={nan} ({time.time()-t0:.2f}s)', flush=True)
    if not ok:
        for (tf, lf), (tp, lp) in zip(fused.info['block_outputs'], plain.info['block_outputs']):
            a, b = fused.tap(tf, B)[..., lf.cols()], plain.tap(tp, B)[..., lp.cols()]
            d = np.abs(a - b)
            print('   stage out', a.shape, 'max|d|', float(d.max()), 'frac bad', float((d > 0).mean()), flush=True)
    fused.close(); plain.close()
    return ok

allok = True
for args in (('shufflenetv2k16', 1, 97, 129), ('shufflenetv2k16', 2, 161, 193), ('shufflenetv2k16', 5, 337, 401),
             ('shufflenetv2k30', 2, 161, 193), ('shufflenetv2k16', 8, 641, 641)):
    allok &= check(*args)
print('DIAG_FUSED ALL OK' if allok else 'DIAG_FUSED MISMATCH', flush=True)

# timing at the bench size
plan = network.random_plan('shufflenetv2k16', seed=0)
B = 64
x = torch.randn(B, 3, 641, 641, generator=torch.Generator().manual_seed(1)).cuda()
for fuse in (False, True):
    net = network.CompiledNet(plan, 641, 641, B, fuse_dw=fuse)
    for _ in range(3):
        net.forward(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        net.forward(x)
    e1.record(); torch.cuda.synchronize()
    ms, kind, flops, nbytes = net.forward_timed(x)
    print(f'TIMING fuse={fuse}: forward {e0.elapsed_time(e1)/10:.3f} ms; by kind: stem {ms[kind==0].sum():.3f} gemm {ms[kind==1].sum():.3f} '
          f'dw {ms[kind==2].sum():.3f} fused {ms[kind==3].sum():.3f}', flush=True)
    for i, o in enumerate(net.op_desc):
        if o['kind'] in ('dw_conv1x1',) or (not fuse and o['kind'] == 'dwconv'):
            hh, ww, _ = net.tensor_shapes[o['in']]
            print(f"   op{i} {o['kind']} {hh}x{ww} ch{o.get('channels')} n{o.get('n_out')} {ms[i]:.4f} ms  {nbytes[i]/ms[i]/1e6:.0f} GB/s", flush=True)
    net.close()
