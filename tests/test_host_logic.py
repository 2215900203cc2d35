"""CPU: host-side mirror of the reference operator surface (statics, params snapshot, pickling)."""
import pickle

import numpy as np
import torch

from openpifpaf_b200 import constants, decoder, synth


def test_static_getset_like_reference():
    C = decoder.CifCaf
    assert C.get_keypoint_threshold() == 0.15 and C.get_reverse_match() is True
    try:
        C.set_force_complete(True)
        C.set_keypoint_threshold(0.0)
        decoder.CifSeeds.set_threshold(0.1)
        decoder.NMSKeypoints.set_instance_threshold(0.0)
        p = C.params()
        assert (p.force_complete, p.keypoint_threshold, p.seed_threshold, p.nms_instance_threshold) == (1, 0.0, 0.1, 0.0)
    finally:
        C.set_force_complete(False)
        C.set_keypoint_threshold(0.15)
        decoder.CifSeeds.set_threshold(0.2)
        decoder.NMSKeypoints.set_instance_threshold(0.15)
    p = C.params()
    assert (p.force_complete, p.keypoint_threshold, p.seed_threshold) == (0, 0.15, 0.2)


def test_pickle_state_is_keypoints_and_skeleton():
    sk = torch.as_tensor(constants.COCO_PERSON_SKELETON, dtype=torch.int64) - 1
    d = decoder.CifCaf(17, sk)
    d2 = pickle.loads(pickle.dumps(d))
    assert d2.n_keypoints == 17 and torch.equal(d2.skeleton, sk)


def test_wholebody_constants_shape():
    sk = constants.wholebody_skeleton()
    assert len(sk) == 160 and max(max(p) for p in sk) == 133 and min(min(p) for p in sk) == 1


def test_synth_is_deterministic_and_in_range():
    a = synth.make_fields('cocokp', 21, 21, 2, 3)
    b = synth.make_fields('cocokp', 21, 21, 2, 3)
    assert synth.fields_digest(a['cif'], a['caf']) == synth.fields_digest(b['cif'], b['caf'])
    assert a['cif'][:, 1].max() <= 1.0 and a['caf'][:, 1].max() <= 1.0 and a['cif'][:, 1].min() >= 0.0
    batch = synth.make_batch('cocokp', 3, 11, 11, 1, seed=1)
    assert batch['cif'].shape == (3, 17, 5, 11, 11) and batch['caf'].shape == (3, 19, 8, 11, 11)
    assert batch['cif'].dtype == np.float32
