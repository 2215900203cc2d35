"""GPU diagnostic: the tcgen05 GEMM (1x1 conv op) in isolation on random data, many (K, N, mode) shapes,
against a float64 numpy product of the same bf16-rounded operands.  Not product code."""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from openpifpaf_b200 import _lib        # noqa: E402

L = _lib.lib()


def bf16_round(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    u = a.view(np.uint32).astype(np.uint64)
    u = ((u + 0x7fff + ((u >> 16) & 1)) >> 16) << 16
    return u.astype(np.uint32).view(np.float32)


def ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def pad8(v):
    return (v + 15) // 16 * 16        # tensors are padded to 16 channels


def run_case(h, w, B, K, N, in_off, shuffle, impl, seed=0):
    rng = np.random.default_rng(seed)
    net = ctypes.c_void_p()
    _lib.check(L.pifpaf_net_create(ctypes.byref(net), 0, B))
    tid = ctypes.c_int32()
    cin_phys = pad8(in_off + K)
    _lib.check(L.pifpaf_net_tensor(net, h, w, cin_phys, ctypes.byref(tid))); t_in = tid.value
    hp = pad8(N)
    wout = pad8(2 * N) if shuffle else hp
    _lib.check(L.pifpaf_net_tensor(net, h, w, wout, ctypes.byref(tid))); t_out = tid.value
    t_src = -1
    if shuffle:
        _lib.check(L.pifpaf_net_tensor(net, h, w, hp, ctypes.byref(tid))); t_src = tid.value
    a = bf16_round(rng.standard_normal((B, h, w, cin_phys)).astype(np.float32))
    wgt = bf16_round((rng.standard_normal((N, K)) / np.sqrt(K)).astype(np.float32))
    bias = rng.standard_normal(N).astype(np.float32)
    _lib.check(L.pifpaf_net_set_tensor(net, t_in, B, ptr(a), a.size))
    src = None
    if shuffle:
        src = bf16_round(rng.standard_normal((B, h, w, hp)).astype(np.float32))
        _lib.check(L.pifpaf_net_set_tensor(net, t_src, B, ptr(src), src.size))
    _lib.check(L.pifpaf_net_conv1x1(net, t_in, in_off, K, N, ptr(wgt), ptr(bias), 1, t_out, 0, t_src, 0))
    _lib.check(L.pifpaf_net_forward(net, None, B, impl, None))
    out = np.empty((B, h, w, wout), dtype=np.float32)
    _lib.check(L.pifpaf_net_tap_tensor(net, t_out, B, ptr(out), out.size))
    ref = np.maximum(a[..., in_off:in_off + K].astype(np.float64) @ wgt.astype(np.float64).T + bias, 0.0)
    if shuffle:
        logical = np.empty(ref.shape[:-1] + (2 * N,))
        logical[..., 0::2] = src[..., :N]
        logical[..., 1::2] = ref
        got, want = out[..., :2 * N], logical
    else:
        got, want = out[..., :N], ref
    err = np.abs(got - want) / (np.abs(want) + 1.0)
    L.pifpaf_net_destroy(net)
    return float(err.max()), got, want


def main():
    ok = True
    cases = [
        # h, w, B, K, N, in_off, shuffle
        (8, 16, 1, 64, 16, 0, False),       # one tile, one k-block, smallest N
        (8, 16, 1, 64, 64, 0, False),
        (8, 16, 1, 128, 64, 0, False),      # two k-blocks
        (8, 16, 1, 24, 176, 0, False),      # K < 64 (stem -> stage2), OOB k fill
        (12, 12, 1, 176, 176, 0, False),    # M = 144: partial second tile; K tail
        (12, 12, 2, 174, 174, 176, False),  # x2 window at an aligned column offset, odd sizes
        (12, 12, 2, 180, 174, 168, False),  # x2 view from floor8(174) (leading pass-through columns)
        (9, 11, 3, 88, 174, 80, True),
        (12, 12, 2, 176, 174, 0, True),     # fused shuffle
        (9, 11, 3, 352, 348, 0, True),      # two n-blocks
        (9, 11, 3, 696, 696, 0, False),     # 4 n-blocks, 11 k-blocks
        (5, 7, 2, 1392, 1392, 0, False),    # conv5
        (41, 41, 2, 1392, 240, 0, False),
        (64, 64, 4, 352, 176, 0, False),    # many tiles per CTA? 16384 rows = 128 tiles
        (161, 161, 2, 176, 174, 0, True),   # 405 tiles > 148 CTAs: persistent loop + accumulator ping-pong
    ]
    for impl, name in ((1, 'simt'), (0, 'tcgen05')):
        for c in cases:
            try:
                err, got, want = run_case(*c, impl)
            except Exception as e:      # noqa: BLE001
                print(f'[{name}] case {c}: EXCEPTION {e}', flush=True)
                ok = False
                continue
            good = err < 2e-2
            ok &= good
            print(f'[{name}] case hwB={c[:3]} K={c[3]} N={c[4]} off={c[5]} shuffle={c[6]}: max rel err {err:.3g} '
                  f'{"ok" if good else "BAD"}', flush=True)
            if not good:
                bad = np.argwhere(np.abs(got - want) / (np.abs(want) + 1.0) > 2e-2)
                rows = (bad[:, 0] * c[0] * c[1] + bad[:, 1] * c[1] + bad[:, 2])
                print('    n bad', len(bad), 'of', got.size, '| rows%128:', sorted(set(rows % 128))[:16], '| tiles:',
                      sorted(set(rows // 128))[:8], '| cols:', sorted(set(bad[:, 3]))[:20])
                print('    sample got/want', got[tuple(bad[0])], want[tuple(bad[0])])
    print('DIAG_GEMM', 'ALL OK' if ok else 'MISMATCH', flush=True)


if __name__ == '__main__':
    main()
