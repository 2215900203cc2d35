"""CPU: the plain-C oracle against the reference's outputs (golden vectors dumped from the
unmodified reference build, and -- when oracle/_ref is present -- the reference itself, live)."""
import hashlib

import numpy as np
import pytest

import helpers
from openpifpaf_b200 import synth
from oracle import cifcaf as oc


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.mark.parametrize('path', helpers.golden_cases(), ids=lambda p: p.split('decoder_')[-1][:-4])
def test_oracle_matches_golden(path):
    g, f, statics, digest_ok = helpers.load_golden(path)
    assert digest_ok, 'synthetic field generator is not bit-reproducible on this machine'
    p = oc.default_params(**helpers.statics_to_params(statics))
    ann, ids, taps = oc.decode(f['cif'], int(g['stride']), f['caf'], int(g['stride']), f['skeleton'],
                               f['n_keypoints'], params=p, taps=True)
    # bit-exact, every stage (the oracle restates libstdc++'s sort/heap tie order too)
    assert sha(taps['cifhr']) == str(g['cifhr_sha256'])
    np.testing.assert_array_equal(taps['seeds_f'], g['seeds_f'])
    np.testing.assert_array_equal(taps['seeds_vxys'], g['seeds_vxys'])
    assert [len(x) for x in taps['fwd']] == list(g['n_fwd'])
    if not statics.get('force_complete'):
        assert sha(np.concatenate([x.reshape(-1, 7) for x in taps['fwd']])) == str(g['fwd_sha256'])
        assert sha(np.concatenate([x.reshape(-1, 7) for x in taps['bwd']])) == str(g['bwd_sha256'])
    np.testing.assert_array_equal(ann, g['annotations'])
    np.testing.assert_array_equal(ids, g['ids'])


def test_golden_stored_fields_roundtrip():
    g = np.load([p for p in helpers.golden_cases() if 'coco11_1p' in p][0])
    ann, ids = oc.decode(g['cif'], 16, g['caf'], 16, synth.make_fields('cocokp', 11, 11, 1, 5)['skeleton'], 17)
    np.testing.assert_array_equal(ann, g['annotations'])


@pytest.mark.parametrize('seed', range(4))
def test_oracle_matches_live_reference(have_reference, seed):
    ref = have_reference
    f = synth.make_fields('cocokp', 21, 27, None, 100 + seed, n_distractors=4)
    ref.ref_configure()
    ra, ri, rt = ref.ref_decode(f['cif'], 16, f['caf'], 16, f['skeleton'], 17, taps=True)
    oa, oi, ot = oc.decode(f['cif'], 16, f['caf'], 16, f['skeleton'], 17, taps=True)
    np.testing.assert_array_equal(rt['cifhr'], ot['cifhr'])
    np.testing.assert_array_equal(rt['seeds_vxys'], ot['seeds_vxys'])
    for a, b in zip(rt['fwd'] + rt['bwd'], ot['fwd'] + ot['bwd']):
        np.testing.assert_array_equal(a, b)
    np.testing.assert_array_equal(ra, oa)
    np.testing.assert_array_equal(ri, oi)


def test_oracle_initial_annotations_match_reference(have_reference):
    ref = have_reference
    f = synth.make_fields('cocokp', 41, 41, 3, 31)
    ref.ref_configure()
    base, _ = ref.ref_decode(f['cif'], 16, f['caf'], 16, f['skeleton'], 17)
    init = base[:1].copy()
    init[0, 5:] = 0.0          # keep a few joints of the first person, let the decoder regrow the rest
    ids = np.array([42], dtype=np.int64)
    ra, ri = ref.ref_decode(f['cif'], 16, f['caf'], 16, f['skeleton'], 17, initial_annotations=init, initial_ids=ids)
    oa, oi = oc.decode(f['cif'], 16, f['caf'], 16, f['skeleton'], 17, initial_annotations=init, initial_ids=ids)
    np.testing.assert_array_equal(ra, oa)
    np.testing.assert_array_equal(ri, oi)
    assert 42 in oi


def test_oracle_grow_connection_blend_matches_reference(have_reference):
    import torch
    have_reference.load_ref()
    rng = np.random.default_rng(0)
    caf = rng.random((50, 7)).astype(np.float32) * np.array([1, 40, 40, 40, 40, 8, 8], dtype=np.float32)
    for only_max in (False, True):
        want = torch.ops.openpifpaf_decoder.grow_connection_blend(torch.from_numpy(caf), 20.0, 20.0, 30.0, 1.0, only_max)
        got = oc.grow_connection_blend(caf, 20.0, 20.0, 30.0, 1.0, only_max)
        assert list(want) == got


def test_empty_and_degenerate_fields():
    sk = synth.make_fields('cocokp', 3, 3, 0, 0)['skeleton']
    cif = np.zeros((17, 5, 3, 3), dtype=np.float32)
    caf = np.zeros((19, 8, 3, 3), dtype=np.float32)
    ann, ids = oc.decode(cif, 16, caf, 16, sk, 17)
    assert ann.shape == (0, 17, 4)
    f = synth.make_fields('cocokp', 1, 1, 0, 0)
    ann, _ = oc.decode(f['cif'], 16, f['caf'], 16, f['skeleton'], 17)
    assert ann.shape[0] == 0


# ---------------------------------------------------------------- CifDet (csrc/src/cifdet.cpp)
@pytest.mark.parametrize('path', helpers.golden_det_cases(), ids=lambda p: p.split('cifdet_')[-1][:-4])
def test_oracle_cifdet_matches_golden(path):
    g, f, digest_ok = helpers.load_golden_det(path)
    assert digest_ok, 'synthetic field generator is not bit-reproducible on this machine'
    cats, scores, boxes = oc.decode_det(f['field'], int(g['stride']))
    np.testing.assert_array_equal(cats, g['categories'])
    np.testing.assert_array_equal(scores, g['scores'])
    np.testing.assert_array_equal(boxes, g['boxes'])
    assert len(cats) <= 120


@pytest.mark.parametrize('seed', range(3))
def test_oracle_cifdet_matches_live_reference(have_reference, seed):
    f = synth.make_det_fields(80, 27, 21, 5 + 20 * seed, 200 + seed, n_distractors=6)
    want = have_reference.ref_decode_det(f['field'], 16)
    got = oc.decode_det(f['field'], 16)
    for a, b in zip(want, got):
        np.testing.assert_array_equal(a, b)
    assert len(got[0]) >= 3


def test_oracle_cifdet_empty_field():
    cats, scores, boxes = oc.decode_det(np.zeros((5, 6, 4, 7), dtype=np.float32), 16)
    assert cats.shape == (0,) and boxes.shape == (0, 4)
