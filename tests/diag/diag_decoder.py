"""GPU diagnostic: CUDA decoder vs the oracle, stage by stage (run under gpurun).
Not a test and not product code; prints one line per case and exits non-zero on mismatch."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from openpifpaf_b200 import synth, decoder as dec   # noqa: E402
from oracle import cifcaf as oc                      # noqa: E402


def compare_case(name, f, stride=16, **cfg):
    p = oc.default_params(seed_sort_stable=1, **cfg)
    oa, oi, ot = oc.decode(f['cif'], stride, f['caf'], stride, f['skeleton'], f['n_keypoints'], params=p, taps=True)
    # configure product statics
    dec.CifCaf.set_force_complete(bool(cfg.get('force_complete', 0)))
    dec.CifCaf.set_greedy(bool(cfg.get('greedy', 0)))
    dec.CifCaf.set_keypoint_threshold(cfg.get('keypoint_threshold', 0.15))
    dec.CifCaf.set_keypoint_threshold_rel(cfg.get('keypoint_threshold_rel', 0.5))
    dec.NMSKeypoints.set_keypoint_threshold(cfg.get('nms_keypoint_threshold', 0.15))
    dec.NMSKeypoints.set_instance_threshold(cfg.get('nms_instance_threshold', 0.15))
    d = dec.CifCaf(f['n_keypoints'], torch.from_numpy(f['skeleton']))
    t0 = time.time()
    ga, gi = d.call(torch.from_numpy(f['cif']), stride, torch.from_numpy(f['caf']), stride)
    t1 = time.time()
    ga, gi = ga.numpy(), gi.numpy()
    hr = d.tap_cifhr().numpy()
    sf, sv = d.tap_seeds()
    fw, bw = d.tap_caf()
    ok_hr = np.array_equal(hr, ot['cifhr'])
    ok_seeds = np.array_equal(sf.numpy(), ot['seeds_f']) and np.array_equal(sv.numpy(), ot['seeds_vxys'])
    ok_caf = all(np.array_equal(a.numpy(), b) for a, b in zip(fw, ot['fwd'])) and \
        all(np.array_equal(a.numpy(), b) for a, b in zip(bw, ot['bwd']))
    ok_n = len(ga) == len(oa)
    maxd = float(np.abs(ga - oa).max()) if ok_n and len(oa) else (0.0 if ok_n else float('nan'))
    exact = ok_n and np.array_equal(ga, oa) and np.array_equal(gi, oi)
    print(f'{name}: N gpu={len(ga)} oracle={len(oa)} | hr {ok_hr} seeds {ok_seeds} (n={len(sf)}/{len(ot["seeds_f"])}) '
          f'caf {ok_caf} ann exact={exact} maxdiff={maxd:.3g} | call {1e3 * (t1 - t0):.1f} ms', flush=True)
    if not ok_hr:
        diff = np.argwhere(hr != ot['cifhr'])
        print('   hr mismatches:', len(diff), 'first', diff[:3].tolist(),
              [(float(hr[tuple(ix)]), float(ot['cifhr'][tuple(ix)])) for ix in diff[:3]])
    if not ok_seeds and len(sf) == len(ot['seeds_f']):
        bad = np.nonzero((sv.numpy() != ot['seeds_vxys']).any(1) | (sf.numpy() != ot['seeds_f']))[0]
        print('   seed mismatches:', len(bad), bad[:5].tolist())
    if not ok_caf:
        for c, (a, b) in enumerate(zip(fw, ot['fwd'])):
            if not np.array_equal(a.numpy(), b):
                print('   caf fwd mismatch at connection', c, a.shape, b.shape)
                break
    return ok_hr and ok_seeds and ok_caf and ok_n and maxd <= 1e-5


def main():
    ok = True
    ok &= compare_case('coco11 1p', synth.make_fields('cocokp', 11, 11, 1, 5))
    for seed in range(3):
        ok &= compare_case(f'coco41 s{seed}', synth.make_fields('cocokp', 41, 41, None, seed, n_distractors=10))
    ok &= compare_case('nonsq 31x41', synth.make_fields('cocokp', 31, 41, 3, 12))
    ok &= compare_case('crowd30', synth.make_fields('cocokp', 41, 41, 30, 7, n_distractors=20))
    ok &= compare_case('greedy', synth.make_fields('cocokp', 41, 41, 4, 13), greedy=1)
    ok &= compare_case('force_complete', synth.make_fields('cocokp', 41, 41, 3, 11, n_distractors=5),
                       force_complete=1, keypoint_threshold=0.0, keypoint_threshold_rel=0.0,
                       nms_keypoint_threshold=0.0, nms_instance_threshold=0.0)
    ok &= compare_case('wholebody 1p', synth.make_fields('wholebody', 41, 41, 1, 21))
    ok &= compare_case('wholebody 4p', synth.make_fields('wholebody', 41, 41, 4, 22))
    print('DIAG', 'ALL OK' if ok else 'MISMATCH', flush=True)
    sys.exit(0 if ok else 1)


if __name__ == '__main__':
    main()
