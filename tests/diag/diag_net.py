"""GPU diagnostic for the network path (run under gpurun): per-op comparison of the CUDA kernels
(tcgen05 and SIMT debug GEMM) against a CPU bf16-rounding emulation of the same op list, then
full-size field parity against the fp32 oracle net and per-op timings.  Not product code."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from openpifpaf_b200 import network                      # noqa: E402
from oracle import net_oracle                             # noqa: E402
import ops_emulator                                       # noqa: E402


def per_op_check(base='shufflenetv2k16', h=97, w=129, B=2):
    shell = net_oracle.make_shell(base, seed=1)
    x = torch.randn(B, 3, h, w)
    plan = network.plan_from_shell(shell)
    tensors, ops, _ = network.build_ops(plan, h, w)
    emu_heads, emu_acts = ops_emulator.run_ops(tensors, ops, x, bf16=True)
    net = network.CompiledNet(plan, h, w, B)
    ok_all = {}
    for impl, name in ((1, 'simt'), (0, 'tcgen05')):
        try:
            heads = net.forward(x.cuda(), gemm_impl=impl)
            torch.cuda.synchronize()
        except Exception as e:      # noqa: BLE001
            print(f'[{name}] forward FAILED: {e}', flush=True)
            ok_all[name] = False
            continue
        worst, first_bad = 0.0, None
        for i, o in enumerate(ops):
            if o['kind'] == 'heads':
                continue
            t = o['out']
            got = net.tap(t, B)
            ref = emu_acts[t].numpy()
            scale = max(float(np.abs(ref).max()), 1e-6)
            err = float(np.abs(got - ref).max()) / scale
            worst = max(worst, err)
            if err > 3e-2 and first_bad is None:
                first_bad = (i, o['kind'], tensors[t], err)
                print(f'[{name}] first bad op #{i} {o["kind"]} out tensor {t} shape {tensors[t]} rel err {err:.3g} '
                      f'(k_cols={o.get("k_cols")}, n_out={o.get("n_out")}, shuffle={o.get("shuffle_src")})', flush=True)
                bad = np.argwhere(np.abs(got - ref) > 3e-2 * scale)
                print('    n bad', len(bad), 'first', bad[:4].tolist(), 'got', got[tuple(bad[0])], 'ref', ref[tuple(bad[0])])
                print('    bad rows (mod 128) sample', sorted(set((bad[:, 1] * tensors[t][1] + bad[:, 2])[:2000] % 128))[:20],
                      'bad cols sample', sorted(set(bad[:, 3][:2000]))[:24])
        herr = []
        for hg, he in zip(heads, emu_heads):
            herr.append(float((hg.cpu() - he).abs().max()))
        ok = first_bad is None and max(herr) < 5e-2
        print(f'[{name}] per-op worst rel err {worst:.3g}; head max abs err vs bf16 emulation {herr} -> {"OK" if ok else "BAD"}',
              flush=True)
        ok_all[name] = ok
    return ok_all


def full_size(base='shufflenetv2k16', size=641, B=8):
    shell = net_oracle.make_shell(base, seed=2)
    x = torch.randn(B, 3, size, size)
    plan = network.plan_from_shell(shell)
    net = network.CompiledNet(plan, size, size, 64)
    xc = x.cuda()
    heads = net.forward(xc)
    torch.cuda.synchronize()
    shell_gpu = shell.cuda()
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    with torch.no_grad():
        ref = shell_gpu(xc)
    for name, hg, hr in zip(('cif', 'caf'), heads, ref):
        d = (hg - hr).abs()
        print(f'full {size}px B={B} {name}: max abs err {float(d.max()):.4g} mean {float(d.mean()):.4g} '
              f'ref std {float(hr.std()):.4g}', flush=True)
    # timing
    for bs in (8, 32, 64):
        xb = torch.randn(bs, 3, size, size, device='cuda')
        for _ in range(3):
            net.forward(xb)
        torch.cuda.synchronize()
        t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
        t0.record()
        n = 5
        for _ in range(n):
            net.forward(xb)
        t1.record(); torch.cuda.synchronize()
        ms = t0.elapsed_time(t1) / n
        print(f'forward bs{bs}: {ms:.2f} ms -> {bs / ms * 1e3:.0f} img/s, '
              f'{net.flops_per_image * bs / ms / 1e9:.1f} TFLOP/s', flush=True)
        if bs == 64:
            ms_op, kind, flops, nbytes = net.forward_timed(xb)
            for k, kn in ((0, 'input_conv'), (1, 'gemm_tc'), (2, 'dwconv')):
                sel = kind == k
                print(f'   {kn}: {sel.sum()} ops {ms_op[sel].sum():.2f} ms, {flops[sel].sum() / ms_op[sel].sum() / 1e9:.1f} TFLOP/s, '
                      f'{nbytes[sel].sum() / ms_op[sel].sum() / 1e6:.0f} GB/s', flush=True)
            views = {}
            for i, o in enumerate(net.op_desc):
                if o['kind'] == 'conv1x1' and o.get('in_off', 0) > 0:
                    hh = net.tensor_shapes[o['out']][0]
                    v = views.setdefault(hh, [o['in_off'], o['k_cols'], 0, 0.0])
                    v[2] += 1; v[3] += float(ms_op[i])
            for hh, (off, kc, cnt, tot) in sorted(views.items(), reverse=True):
                print(f'   chunk-view GEMMs at {hh}x{hh}: start column {off}, k_cols {kc}: {cnt} ops {tot:.3f} ms '
                      f'({tot / cnt * 1e3:.1f} us each)', flush=True)
            if os.environ.get('DIAG_NET_FAST'):
                return
            order = np.argsort(-ms_op)[:8]
            for i in order:
                print(f'      op {i} kind {kind[i]} {ms_op[i]:.3f} ms {flops[i] / ms_op[i] / 1e9:.1f} TFLOP/s '
                      f'{nbytes[i] / ms_op[i] / 1e6:.0f} GB/s')
            # stock PyTorch (cuDNN) bar: same Shell, fp32 / tf32 / bf16
            with torch.no_grad():
                for label, prep in (('fp32', None), ('bf16 channels_last', 'bf16')):
                    m, xx = shell_gpu, xb
                    if prep == 'bf16':
                        m = shell_gpu.to(torch.bfloat16).to(memory_format=torch.channels_last)
                        xx = xb.to(torch.bfloat16).to(memory_format=torch.channels_last)
                    for _ in range(2):
                        m(xx)
                    torch.cuda.synchronize()
                    t0.record()
                    for _ in range(3):
                        m(xx)
                    t1.record(); torch.cuda.synchronize()
                    msr = t0.elapsed_time(t1) / 3
                    print(f'   stock torch {label} forward bs64: {msr:.2f} ms -> {64 / msr * 1e3:.0f} img/s', flush=True)
                shell_gpu.float()


def main():
    ok = per_op_check()
    try:
        full_size()
    except Exception as e:      # noqa: BLE001
        import traceback
        traceback.print_exc()
        print('full_size FAILED', e)
    print('DIAG_NET', ok, flush=True)


if __name__ == '__main__':
    main()
