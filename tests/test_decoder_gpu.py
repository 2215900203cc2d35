"""GPU: parity of the CUDA CifCaf decoder (through the C ABI) with the oracle / golden vectors."""
import numpy as np
import pytest
import torch

import helpers
from openpifpaf_b200 import decoder, synth
from oracle import cifcaf as oc

pytestmark = pytest.mark.gpu


def configure(statics):
    C = decoder.CifCaf
    C.set_greedy(bool(statics.get('greedy', False)))
    C.set_force_complete(bool(statics.get('force_complete', False)))
    C.set_keypoint_threshold(statics.get('keypoint_threshold', 0.15))
    C.set_keypoint_threshold_rel(statics.get('keypoint_threshold_rel', 0.5))
    decoder.NMSKeypoints.set_keypoint_threshold(statics.get('nms_keypoint_threshold', 0.15))
    decoder.NMSKeypoints.set_instance_threshold(statics.get('nms_instance_threshold', 0.15))


@pytest.fixture(autouse=True)
def reset_statics():
    yield
    configure({})


def make_decoder(f):
    return decoder.CifCaf(f['n_keypoints'], torch.from_numpy(f['skeleton']))


@pytest.mark.parametrize('path', helpers.golden_cases(), ids=lambda p: p.split('decoder_')[-1][:-4])
def test_cuda_matches_reference_golden(path):
    """CUDA vs the outputs the unmodified reference produced (tests/golden)."""
    g, f, statics, digest_ok = helpers.load_golden(path)
    assert digest_ok
    configure(statics)
    d = make_decoder(f)
    stride = int(g['stride'])
    ann, ids = d.call(torch.from_numpy(f['cif']), stride, torch.from_numpy(f['caf']), stride)
    helpers.assert_annotations_close(ann.numpy(), g['annotations'], 'annotations')
    np.testing.assert_array_equal(ids.numpy(), g['ids'])
    sf, sv = d.tap_seeds()
    assert len(sf) == len(g['seeds_f'])
    # seed order may differ from std::sort only inside runs of exactly tied scores
    np.testing.assert_array_equal(sv.numpy()[:, 0], g['seeds_vxys'][:, 0])
    if not statics.get('force_complete'):
        fw, bw = d.tap_caf()
        assert [len(x) for x in fw] == list(g['n_fwd']) and [len(x) for x in bw] == list(g['n_bwd'])


@pytest.mark.parametrize('seed', range(6))
@pytest.mark.parametrize('shape', [(41, 41), (23, 37)])
def test_cuda_matches_oracle_all_stages(seed, shape):
    """bit-exact CifHr map / seeds / CAF lists and in-tolerance annotations on seeded inputs."""
    h, w = shape
    f = synth.make_fields('cocokp', h, w, None, 500 + seed, n_distractors=6)
    p = oc.default_params(seed_sort_stable=1)
    oa, oi, ot = oc.decode(f['cif'], 16, f['caf'], 16, f['skeleton'], 17, params=p, taps=True)
    d = make_decoder(f)
    ga, gi = d.call(torch.from_numpy(f['cif']), 16, torch.from_numpy(f['caf']), 16)
    np.testing.assert_array_equal(d.tap_cifhr().numpy(), ot['cifhr'])
    sf, sv = d.tap_seeds()
    np.testing.assert_array_equal(sf.numpy(), ot['seeds_f'])
    np.testing.assert_array_equal(sv.numpy(), ot['seeds_vxys'])
    fw, bw = d.tap_caf()
    for a, b in zip(fw + bw, ot['fwd'] + ot['bwd']):
        np.testing.assert_array_equal(a.numpy(), b)
    helpers.assert_annotations_close(ga.numpy(), oa, 'annotations')
    np.testing.assert_array_equal(gi.numpy(), oi)


def test_get_cifhr_like_reference():
    f = synth.make_fields('cocokp', 11, 11, 1, 5)
    d = make_decoder(f)
    d.call(torch.from_numpy(f['cif']), 16, torch.from_numpy(f['caf']), 16)
    hr, rev = d.get_cifhr()
    assert rev == 1.0 and tuple(hr.shape) == (17, 161, 161)
    assert float(hr.max()) <= 2.0 and float(hr[hr > 0].min()) >= 1.0     # stored as revision + value


def test_initial_annotations():
    f = synth.make_fields('cocokp', 41, 41, 3, 31)
    base, _ = oc.decode(f['cif'], 16, f['caf'], 16, f['skeleton'], 17)
    init = base[:1].copy()
    init[0, 5:] = 0.0
    ids = np.array([42], dtype=np.int64)
    oa, oi = oc.decode(f['cif'], 16, f['caf'], 16, f['skeleton'], 17, initial_annotations=init, initial_ids=ids)
    d = make_decoder(f)
    ga, gi = d.call_with_initial_annotations(torch.from_numpy(f['cif']), 16, torch.from_numpy(f['caf']), 16,
                                             torch.from_numpy(init), torch.from_numpy(ids))
    helpers.assert_annotations_close(ga.numpy(), oa, 'initial annotations')
    np.testing.assert_array_equal(gi.numpy(), oi)


def test_grow_connection_blend_free_op():
    rng = np.random.default_rng(0)
    caf = rng.random((300, 7)).astype(np.float32) * np.array([1, 40, 40, 40, 40, 8, 8], dtype=np.float32)
    for only_max in (False, True):
        for (x, y, s) in ((20.0, 20.0, 30.0), (5.0, 33.0, 12.0), (100.0, 100.0, 2.0)):
            want = oc.grow_connection_blend(caf, x, y, s, 1.0, only_max)
            got = decoder.grow_connection_blend(torch.from_numpy(caf), x, y, s, 1.0, only_max)
            np.testing.assert_allclose(got, want, rtol=0, atol=1e-6)


def test_edge_cases_empty_tiny_and_ragged():
    sk = synth.make_fields('cocokp', 3, 3, 0, 0)['skeleton']
    d = decoder.CifCaf(17, torch.from_numpy(sk))
    ann, ids = d.call(torch.zeros(17, 5, 3, 3), 16, torch.zeros(19, 8, 3, 3), 16)
    assert tuple(ann.shape) == (0, 17, 4) and tuple(ids.shape) == (0,)
    f = synth.make_fields('cocokp', 1, 1, 0, 0)            # 1x1 field, hi-res map 1x1
    ann, _ = d.call(torch.from_numpy(f['cif']), 16, torch.from_numpy(f['caf']), 16)
    assert ann.shape[0] == 0
    # ragged: different shapes through the same handle, larger first then smaller, then non-square
    for (h, w, n) in ((41, 41, 2), (11, 11, 1), (17, 29, 1)):
        f = synth.make_fields('cocokp', h, w, n, 3)
        oa, _ = oc.decode(f['cif'], 16, f['caf'], 16, f['skeleton'], 17, params=oc.default_params(seed_sort_stable=1))
        ga, _ = d.call(torch.from_numpy(f['cif']), 16, torch.from_numpy(f['caf']), 16)
        helpers.assert_annotations_close(ga.numpy(), oa, f'{h}x{w}')


def test_capacity_overflow_is_an_error():
    f = synth.make_fields('cocokp', 41, 41, 6, 2, n_distractors=30)
    d = make_decoder(f)
    d.max_annotations = 4
    with pytest.raises(RuntimeError, match='capacity'):
        d.call(torch.from_numpy(f['cif']), 16, torch.from_numpy(f['caf']), 16)


def test_batched_device_path_full_size_properties():
    """BASELINE full size (bs 64, 41x41 cells): every image of a batched device-resident decode equals its
    single-image decode (images are independent), a repeated decode is idempotent (epoch-tagged occupancy,
    no stale state), and a sample of images is checked against the oracle."""
    B = 64
    batch = synth.make_batch('cocokp', B, 41, 41, None, seed=3)
    sk = torch.from_numpy(batch['skeleton'])
    d = decoder.CifCaf(17, sk)
    cif = torch.from_numpy(batch['cif']).cuda()
    caf = torch.from_numpy(batch['caf']).cuda()
    r1 = d.decode_batch(cif, 16, caf, 16)
    r2 = d.decode_batch(cif, 16, caf, 16)
    assert len(r1) == B
    for (a1, i1), (a2, i2) in zip(r1, r2):
        assert torch.equal(a1, a2) and torch.equal(i1, i2)
    single = decoder.CifCaf(17, sk)
    p = oc.default_params(seed_sort_stable=1)
    n_total = 0
    for b in range(B):
        sa, _ = single.call(torch.from_numpy(batch['cif'][b]), 16, torch.from_numpy(batch['caf'][b]), 16)
        assert torch.equal(sa, r1[b][0]), f'image {b}: batched != single'
        n_total += len(sa)
        if b % 8 == 0:
            oa, _ = oc.decode(batch['cif'][b], 16, batch['caf'][b], 16, batch['skeleton'], 17, params=p)
            helpers.assert_annotations_close(sa.numpy(), oa, f'image {b}')
    assert n_total == sum(batch['n_planted'])      # every planted person decoded, nothing else


def test_wholebody_wide_caf_batch():
    """configs[2]: 133 keypoints / 160 associations."""
    batch = synth.make_batch('wholebody', 2, 41, 41, 2, seed=5)
    d = decoder.CifCaf(133, torch.from_numpy(batch['skeleton']))
    res = d.decode_batch(torch.from_numpy(batch['cif']).cuda(), 16, torch.from_numpy(batch['caf']).cuda(), 16)
    p = oc.default_params(seed_sort_stable=1)
    for b in range(2):
        oa, _ = oc.decode(batch['cif'][b], 16, batch['caf'][b], 16, batch['skeleton'], 133, params=p)
        helpers.assert_annotations_close(res[b][0].numpy(), oa, f'wholebody image {b}')
        assert len(oa) == 2


def test_three_hundred_decodes_on_one_handle_cross_every_tag_wrap():
    """State-machine edges of a long-lived handle (a Predictor keeps ONE for its lifetime): the byte-wide occupancy
    tags wrap every 127 decodes (real clear of the map), the 32-bit CifHr tile epoch wraps once in 2^32 decodes
    (moved next to the wrap through the debug entry point).  Every decode must equal the first, bit for bit --
    annotations, seeds and the CifHr map -- and pipelined fetches (two outstanding) must stay in order."""
    from openpifpaf_b200 import _lib
    fa = synth.make_fields('cocokp', 17, 21, 3, 901, n_distractors=4)
    fb = synth.make_fields('cocokp', 17, 21, 2, 902, n_distractors=4)
    d = make_decoder(fa)
    cif = torch.from_numpy(np.stack([fa['cif'], fb['cif']])).cuda()
    caf = torch.from_numpy(np.stack([fa['caf'], fb['caf']])).cuda()
    first = d.decode_batch(cif, 16, caf, 16)
    hr0 = d.tap_cifhr(0).clone()
    p = oc.default_params(seed_sort_stable=1)
    for b, f in enumerate((fa, fb)):
        oa, _ = oc.decode(f['cif'], 16, f['caf'], 16, f['skeleton'], 17, params=p)
        helpers.assert_annotations_close(first[b][0].numpy(), oa, f'image {b}')
        assert len(oa) >= 2
    # CifHr epoch: 3 decodes before the 32-bit wrap
    _lib.check(_lib.lib().pifpaf_decoder_debug_set_epochs(d._handle, 1, 0xFFFFFFFD))
    for i in range(300):
        if i % 3 == 2:          # pipelined: two decodes in flight, fetched in order
            d.decode_batch_async(cif, 16, caf, 16)
            d.fetch_begin()
            d.decode_batch_async(cif.flip(0).contiguous(), 16, caf.flip(0).contiguous(), 16)
            d.fetch_begin()
            r1, r2 = d.fetch_end(), d.fetch_end()
            for b in range(2):
                assert torch.equal(r1[b][0], first[b][0]) and torch.equal(r2[1 - b][0], first[b][0]), i
        else:
            res = d.decode_batch(cif, 16, caf, 16)
            for b in range(2):
                assert torch.equal(res[b][0], first[b][0]) and torch.equal(res[b][1], first[b][1]), i
        if i in (1, 2, 3, 4, 126, 127, 128, 254, 299):
            # the pipelined branch decoded the flipped batch last: image fa sits at index 1 there
            assert torch.equal(d.tap_cifhr(1 if i % 3 == 2 else 0), hr0), i
