"""Generate tests/golden/*.npz from the UNMODIFIED reference decoder (oracle/_ref).

Run in the build container (needs /root/reference to build oracle/_ref):
    python -m oracle.make_golden
Each fixture stores the generator arguments, the sha256 of the generated float32 fields
(openpifpaf_b200.synth is bit-reproducible, so inputs need not be stored), and the reference's
outputs on a FRESH CifCaf instance: annotations, ids, sorted seeds, per-connection CafScored
counts and a sha256 of the CifHr map.  The smallest case also stores the raw fields.
TEST INFRASTRUCTURE."""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from openpifpaf_b200 import synth      # noqa: E402
from oracle import cifcaf as oc        # noqa: E402

CASES = [
    # name, workload, h, w, n_people, seed, n_distractors, stride, reference statics
    ('coco11_1p', 'cocokp', 11, 11, 1, 5, 0, 16, {}),
    ('coco11_2p_d3', 'cocokp', 11, 11, 2, 6, 3, 16, {}),
    ('coco41_poisson_s0', 'cocokp', 41, 41, None, 0, 10, 16, {}),
    ('coco41_poisson_s1', 'cocokp', 41, 41, None, 1, 10, 16, {}),
    ('coco31x41_3p', 'cocokp', 31, 41, 3, 12, 0, 16, {}),
    ('coco51_5p_stride8', 'cocokp', 51, 51, 5, 14, 4, 8, {}),
    ('crowd30', 'cocokp', 41, 41, 30, 7, 20, 16, {}),
    ('greedy_4p', 'cocokp', 41, 41, 4, 13, 0, 16, {'greedy': True}),
    ('force_complete_3p', 'cocokp', 41, 41, 3, 11, 5, 16,
     {'force_complete': True, 'keypoint_threshold': 0.0, 'keypoint_threshold_rel': 0.0,
      'nms_keypoint_threshold': 0.0, 'nms_instance_threshold': 0.0}),
    ('wholebody_1p', 'wholebody', 41, 41, 1, 21, 0, 16, {}),
    ('wholebody_4p', 'wholebody', 41, 41, 4, 22, 0, 16, {}),
]


DET_CASES = [
    # name, n_categories, h, w, n_objects, seed, n_distractors, stride
    ('det11_3cat_1', 3, 11, 11, 1, 2, 2, 16),
    ('det41_80cat_6', 80, 41, 41, 6, 0, 4, 16),
    ('det21x33_80cat_3', 80, 21, 33, 3, 1, 4, 16),
    ('det51_80cat_150', 80, 51, 51, 150, 4, 10, 8),         # hits max_detections_before_nms = 120
]


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    out_dir = os.path.join(ROOT, 'tests', 'golden')
    os.makedirs(out_dir, exist_ok=True)
    for name, workload, h, w, n_people, seed, n_dis, stride, statics in CASES:
        f = synth.make_fields(workload, h, w, n_people, seed, n_dis)
        oc.ref_configure(**statics)
        ann, ids, taps = oc.ref_decode(f['cif'], stride, f['caf'], stride, f['skeleton'], f['n_keypoints'], taps=True)
        data = dict(
            workload=workload, h=h, w=w, n_people=-1 if n_people is None else n_people, seed=seed,
            n_distractors=n_dis, stride=stride,
            statics_keys=np.array(list(statics.keys()), dtype='U64'),
            statics_vals=np.array([float(v) for v in statics.values()], dtype=np.float64),
            fields_sha256=synth.fields_digest(f['cif'], f['caf']),
            annotations=ann, ids=ids,
            seeds_f=taps['seeds_f'], seeds_vxys=taps['seeds_vxys'],
            n_fwd=np.array([len(x) for x in taps['fwd']], dtype=np.int64),
            n_bwd=np.array([len(x) for x in taps['bwd']], dtype=np.int64),
            fwd_sha256=sha(np.concatenate([x.reshape(-1, 7) for x in taps['fwd']])),
            bwd_sha256=sha(np.concatenate([x.reshape(-1, 7) for x in taps['bwd']])),
            cifhr_sha256=sha(taps['cifhr']), cifhr_sum=float(taps['cifhr'].astype(np.float64).sum()),
        )
        if name == 'coco11_1p':
            data['cif'] = f['cif']
            data['caf'] = f['caf']
        path = os.path.join(out_dir, f'decoder_{name}.npz')
        np.savez_compressed(path, **data)
        print(name, 'N =', len(ann), 'seeds =', len(taps['seeds_f']), os.path.getsize(path), 'bytes')
    oc.ref_configure()
    # CifDet (csrc/src/cifdet.cpp): raw output of a FRESH torch.classes.openpifpaf_decoder.CifDet instance
    for name, n_cat, h, w, n_obj, seed, n_dis, stride in DET_CASES:
        f = synth.make_det_fields(n_cat, h, w, n_obj, seed, n_dis)
        cats, scores, boxes = oc.ref_decode_det(f['field'], stride)
        path = os.path.join(out_dir, f'cifdet_{name}.npz')
        np.savez_compressed(path, n_categories=n_cat, h=h, w=w, n_objects=n_obj, seed=seed, n_distractors=n_dis,
                            stride=stride, field_sha256=sha(f['field']), categories=cats, scores=scores, boxes=boxes)
        print(name, 'N =', len(cats), os.path.getsize(path), 'bytes')


if __name__ == '__main__':
    main()
