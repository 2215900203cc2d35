"""GPU: decoder-level agreement between the fp32 network and the bf16 CUDA network ("report max-abs/rel
error on fields and decoder-level agreement on counts").

No trained checkpoint can be downloaded, and a random backbone's feature map is almost rank one over positions (a
closed-form fit of the heads to pose targets needs weights of norm ~1e4, measured), so poses are planted where the
network is linear: a per-position offset D is added to the conv5 / block5 feature map such that the fp32 heads emit
the planted fields EXACTLY,  W (f + D) + b = T  (D = pinv(W) (T - W f - b), minimum norm, float64).  Then

  fp32 path : fp32 PyTorch backbone (no TF32)  -> f   ; fp32 heads on f  + D  -> fields32 (== planted poses)
  bf16 path : CUDA bf16 backbone (this repo)   -> f~  ; CUDA heads GEMM on bf16(f~ + D), bf16 weights -> fields16

so fields16 carries every error the product path makes on such an input: the backbone's accumulated bf16 error pushed
through the head weights, the bf16 rounding of the head input and weights, the f32 epilogue.  Both field sets are
decoded by the same CUDA decoder; instance counts must be identical and the keypoint / score deltas are reported
(written to gpurun_out/bf16_accuracy.json; the committed copy is profiles/r2_bf16_accuracy.json)."""
import json
import os

import numpy as np
import pytest
import torch

from openpifpaf_b200 import constants, decoder, network, synth
from oracle import net_oracle

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# tolerances of the bf16 path against fp32 at the decoder level, in input-image pixels / score units
# (CPU dry run with the bf16 emulator of tests/ops_emulator.py, k16 at 321 px: 1.3 px, 0.13 -- the backbone's bf16
# error accumulated over 58 layers is ~1.8 % of the feature range and reaches the ~N(0,1) head pre-activations as
# noise of sigma ~0.03, max ~0.13)
# measured on B200 (profiles/r2_bf16_accuracy.json): k16 641 px 1.30 px / 0.16 / 8 of 434 joints present in one
# decode only (all next to the 0.15 keypoint threshold); resnet50 1.11 px / 0.12; k30-wholebody (100 bf16 layers,
# joint scales 0.15 x person: narrow Gaussians, scores react strongly to sub-pixel shifts) 1.51 px / 0.38
XY_TOL_PX = 4.0
SCORE_TOL = {'cocokp': 0.25, 'wholebody': 0.5}
JOINT_MISMATCH_FRAC = 0.05


def _logit(p):
    p = np.clip(p, 1e-4, 1.0 - 1e-4)
    return np.log(p / (1.0 - p))


def _inv_softplus(s):
    return np.log(np.expm1(np.maximum(s, 1e-3)))


def _preactivations(fields, n_vectors):
    """inverse of the CompositeField4 eval epilogue (heads.py:360-378): [B,F,comp,h,w] fields -> pre-activations"""
    b, f, ncomp, h, w = fields.shape
    ii, jj = np.meshgrid(np.arange(w, dtype=np.float64), np.arange(h, dtype=np.float64))
    out = np.zeros(fields.shape, dtype=np.float64)
    out[:, :, 1] = _logit(fields[:, :, 1].astype(np.float64))
    for v in range(n_vectors):
        out[:, :, 2 + 2 * v] = fields[:, :, 2 + 2 * v] - ii
        out[:, :, 3 + 2 * v] = fields[:, :, 3 + 2 * v] - jj
    for c in range(2 + 2 * n_vectors, ncomp):
        out[:, :, c] = _inv_softplus(fields[:, :, c].astype(np.float64))
    return out


def _calibrate(shell, feat, seed):
    """random heads with ~N(0,1) pre-activations over positions and a sparse-confidence bias (what
    network.calibrate_random_heads does for the bench), on the torch Shell"""
    g = torch.Generator().manual_seed(seed)
    f = feat.permute(0, 2, 3, 1).reshape(-1, feat.shape[1]).double().cpu()
    mu = f.mean(0)
    std = float((f - mu).pow(2).mean().sqrt())
    with torch.no_grad():
        for hn in shell.head_nets:
            m = hn.meta
            ncomp = 1 + m.n_confidences + 2 * m.n_vectors + m.n_scales
            w = torch.randn(hn.conv.weight.shape[:2], generator=g).double() / (feat.shape[1] ** 0.5 * max(std, 1e-9))
            b = torch.randn(w.shape[0], generator=g).double() * 0.1 - w @ mu
            b.view(m.n_fields, ncomp)[:, 1:1 + m.n_confidences] -= 2.5
            hn.conv.weight.copy_(w.float().view_as(hn.conv.weight))
            hn.conv.bias.copy_(b.float())


def _match(a, b):
    """greedy nearest matching of instances by mean keypoint distance (orders may differ when scores are close)"""
    used, pairs = set(), []
    for i in range(a.shape[0]):
        best, bj = None, -1
        for j in range(b.shape[0]):
            if j in used:
                continue
            both = (a[i, :, 0] > 0) & (b[j, :, 0] > 0)
            d = float(np.abs(a[i, both, 1:3] - b[j, both, 1:3]).mean()) if both.any() else 1e9
            if best is None or d < best:
                best, bj = d, j
        used.add(bj)
        pairs.append((i, bj))
    return pairs


@pytest.mark.parametrize('base,workload,size,batch,n_people', [
    ('shufflenetv2k16', 'cocokp', 641, 4, None),          # BASELINE C0/C2 network at size, Poisson(4)+1 people
    ('resnet50', 'cocokp', 321, 2, 3),                    # C4 family
    ('resnet50', 'cocokp', 801, 2, 4),                    # BASELINE C4 at size (801 px, 51 x 51 cells)
    ('shufflenetv2k30', 'wholebody', 641, 2, 2),          # C3 at size: 133 keypoints / 160 connections
])
def test_bf16_network_decodes_like_fp32(base, workload, size, batch, n_people):
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    K = synth.WORKLOADS[workload][0]
    skeleton1 = np.asarray(synth.skeleton_for(workload), dtype=np.int64)
    shell = net_oracle.make_shell(base, n_keypoints=K, n_connections=skeleton1.shape[0], seed=5, he_init=True)
    x = torch.randn(batch, 3, size, size, generator=torch.Generator().manual_seed(9))
    h = w = (size - 1) // 16 + 1
    with torch.no_grad():
        feat32 = shell.cuda().base_net(x.cuda())                       # [B,C,h,w] fp32, no TF32
    assert tuple(feat32.shape[2:]) == (h, w)
    _calibrate(shell, feat32, seed=6)
    shell = shell.cuda()

    planted = synth.make_batch(workload, batch, h, w, n_people, seed=21)
    targets = np.concatenate([_preactivations(planted['cif'], 1).transpose(0, 3, 4, 1, 2).reshape(batch * h * w, -1),
                              _preactivations(planted['caf'], 2).transpose(0, 3, 4, 1, 2).reshape(batch * h * w, -1)], axis=1)
    W = torch.cat([hn.conv.weight.reshape(hn.conv.weight.shape[0], -1) for hn in shell.head_nets]).detach().double().cpu().numpy()
    bias = torch.cat([hn.conv.bias for hn in shell.head_nets]).detach().double().cpu().numpy()
    F32 = feat32.permute(0, 2, 3, 1).reshape(batch * h * w, -1).double().cpu().numpy()
    resid = targets - (F32 @ W.T + bias)
    D = resid @ np.linalg.pinv(W).T                                     # [P, C] minimum-norm offsets
    D4 = D.reshape(batch, h, w, -1)

    # ---- fp32 path
    with torch.no_grad():
        f_plus = feat32 + torch.from_numpy(D4).permute(0, 3, 1, 2).float().cuda()
        fields32 = [hn(f_plus).contiguous() for hn in shell.head_nets]
    # sanity: the fp32 network does emit the planted poses
    assert float((fields32[0][:, :, 1].cpu() - torch.from_numpy(planted['cif'][:, :, 1])).abs().max()) < 2e-2

    # ---- bf16 path through the product kernels
    plan = network.plan_from_shell(shell.cpu())
    net = network.CompiledNet(plan, size, size, batch)
    net.forward(x.cuda())
    torch.cuda.synchronize()
    t_feat, lay = net.info['feature']
    f_tilde = net.tap(t_feat, batch)[..., lay.cols()].astype(np.float64)                  # bf16 values, exact in f32
    heads_net = network.CompiledNet(network.heads_only_plan(plan), h, w, batch)
    fields16 = [t.clone() for t in heads_net.forward_features((f_tilde + D4).astype(np.float32))]

    # ---- field-level error (reported), decoder-level agreement (asserted)
    dec = decoder.CifCaf(K, torch.from_numpy(skeleton1 - 1))
    res32 = dec.decode_batch(fields32[0], 16, fields32[1], 16)
    res16 = dec.decode_batch(fields16[0], 16, fields16[1], 16)
    counts32 = [int(a.shape[0]) for a, _ in res32]
    counts16 = [int(a.shape[0]) for a, _ in res16]
    dxy = dv = 0.0
    n_kp = n_mismatch = 0
    for (a32, _), (a16, _) in zip(res32, res16):
        a32, a16 = a32.numpy(), a16.numpy()
        if a32.shape != a16.shape:
            continue
        for i, j in _match(a32, a16):
            both = (a32[i, :, 0] > 0) & (a16[j, :, 0] > 0)
            n_mismatch += int(((a32[i, :, 0] > 0) != (a16[j, :, 0] > 0)).sum())
            n_kp += int(both.sum())
            if both.any():
                dxy = max(dxy, float(np.abs(a32[i, both, 1:3] - a16[j, both, 1:3]).max()))
                dv = max(dv, float(np.abs(a32[i, both, 0] - a16[j, both, 0]).max()))
    report = {
        'network': base, 'workload': workload, 'input': [batch, 3, size, size],
        'planted_people': [int(n) for n in planted['n_planted']],
        'instances_fp32': counts32, 'instances_bf16': counts16, 'keypoints_compared': n_kp,
        'keypoints_present_in_one_only': n_mismatch,
        'max_abs_dxy_px': round(dxy, 5), 'max_abs_dscore': round(dv, 6),
        'field_max_abs_err': [round(float((a - b).abs().max()), 5) for a, b in zip(fields32, fields16)],
        'field_err_over_std': [round(float((a - b).abs().max() / a.std()), 5) for a, b in zip(fields32, fields16)],
        'feature_rel_err': round(float(np.abs(f_tilde - F32.reshape(f_tilde.shape)).max() / np.abs(F32).max()), 5),
    }
    out_dir = os.path.join(ROOT, 'gpurun_out')
    os.makedirs(out_dir, exist_ok=True)
    path = os.path.join(out_dir, 'bf16_accuracy.json')
    prev = json.load(open(path)) if os.path.exists(path) else {}
    prev[f'{base}:{workload}:{size}'] = report
    with open(path, 'w') as f:
        json.dump(prev, f, indent=1)
    print(json.dumps(report))
    assert counts32 == counts16, report
    assert sum(counts32) >= sum(planted['n_planted']) * 0.8, report       # the planted people are found
    assert n_mismatch <= JOINT_MISMATCH_FRAC * max(n_kp, 1), report
    assert dxy <= XY_TOL_PX and dv <= SCORE_TOL[workload], report
