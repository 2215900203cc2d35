"""GPU diagnostic: decoder-only timing on planted-pose fields (device-resident, batched).  Not product code."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from openpifpaf_b200 import decoder, synth      # noqa: E402


def run(workload, B, n_people, reps=20, label=''):
    batch = synth.make_batch(workload, B, 41, 41, n_people, seed=11)
    K = batch['n_keypoints']
    d = decoder.CifCaf(K, torch.from_numpy(batch['skeleton']))
    cif = torch.from_numpy(batch['cif']).cuda()
    caf = torch.from_numpy(batch['caf']).cuda()
    for _ in range(3):
        res = d.decode_batch(cif, 16, caf, 16)
    n_ann = sum(len(a) for a, _ in res)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        d.decode_batch_async(cif, 16, caf, 16)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    st = d.last_stats()
    ck = st['grow_clocks']
    print(f'   k_grow per image: rounds {st["grow_rounds"] / B:.2f}, seeds grown {st["grow_seeds_grown"] / B:.1f}, '
          f'annotations {st["annotations_before_nms"] / B:.1f}; kclocks setup {ck["setup"] / B / 1e3:.0f} select {ck["select"] / B / 1e3:.0f} '
          f'grow {ck["grow"] / B / 1e3:.0f} commit {ck["commit"] / B / 1e3:.0f}', flush=True)
    print(f'{label or workload} B={B} people={n_people}: {ms:.3f} ms/batch = {ms / B:.4f} ms/img, '
          f'{n_ann} annotations (planted {sum(batch["n_planted"])})', flush=True)


if __name__ == '__main__':
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    run('cocokp', 64, None, reps, 'coco poisson(4)+1')
    run('cocokp', 1, 5, reps, 'coco single image 5 people')
    run('cocokp', 8, 30, reps, 'crowd (config 5)')
    run('wholebody', 16, 3, max(reps // 4, 2), 'wholebody (config 3)')
