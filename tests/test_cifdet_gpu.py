"""GPU: the CUDA CifDet decoder against the golden vectors of the unmodified reference
(csrc/src/cifdet.cpp through oracle/_ref), the plain-C oracle, and torchvision's NMS as the reference's Python
wrapper applies it (decoder/cifdet.py:55-64).  Integer / index outputs bit-exact; scores and boxes bit-exact too
(same float arithmetic, -fmad=false)."""
import numpy as np
import pytest
import torch

import helpers
from openpifpaf_b200 import decoder, synth
from oracle import cifcaf as oc

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('path', helpers.golden_det_cases(), ids=lambda p: p.split('cifdet_')[-1][:-4])
def test_cifdet_matches_reference_golden(path):
    g, f, digest_ok = helpers.load_golden_det(path)
    assert digest_ok
    d = decoder.CifDet()
    cats, scores, boxes = d.call(torch.from_numpy(f['field']), int(g['stride']))
    np.testing.assert_array_equal(cats.numpy(), g['categories'])
    np.testing.assert_array_equal(scores.numpy(), g['scores'])
    np.testing.assert_array_equal(boxes.numpy(), g['boxes'])


def test_cifdet_batched_device_fields_match_oracle():
    """ragged content, non-square fields, one handle, batch of 5; also after 300 decodes on the same handle (the
    byte-wide occupancy tags wrap)"""
    fields = [synth.make_det_fields(80, 23, 37, n, 300 + n, n_distractors=5)['field'] for n in (0, 1, 7, 40, 130)]
    d = decoder.CifDet()
    dev = torch.from_numpy(np.stack(fields)).cuda()
    p = oc.default_params(seed_sort_stable=1)
    want = [oc.decode_det(f, 8, params=p) for f in fields]
    for rep in range(3):
        got = d.decode_batch(dev, 8)
        for (gc, gs, gb), (wc, ws, wb) in zip(got, want):
            np.testing.assert_array_equal(gc.numpy(), wc)
            np.testing.assert_array_equal(gs.numpy(), ws)
            np.testing.assert_array_equal(gb.numpy(), wb)
        if rep == 0:
            for _ in range(300):
                d.decode_batch(dev[:2], 8)
    assert len(want[4][0]) == 120 and len(want[0][0]) <= 8


def test_cifdet_statics_like_reference():
    """CifDet.max_detections_before_nms / CifDetSeeds.threshold / CifHr.threshold (module.cpp:57-62, 96-97)"""
    f = synth.make_det_fields(80, 31, 31, 60, 5, n_distractors=5)['field']
    d = decoder.CifDet()
    try:
        decoder.CifDet.set_max_detections_before_nms(17)
        decoder.CifDetSeeds.set_threshold(0.4)
        cats, scores, boxes = d.call(torch.from_numpy(f), 16)
        wc, ws, wb = oc.decode_det(f, 16, params=oc.default_params(seed_sort_stable=1, seed_threshold=0.4),
                                   max_detections_before_nms=17)
        assert len(wc) == 17
        np.testing.assert_array_equal(cats.numpy(), wc)
        np.testing.assert_array_equal(scores.numpy(), ws)
        np.testing.assert_array_equal(boxes.numpy(), wb)
    finally:
        decoder.CifDet.set_max_detections_before_nms(120)
        decoder.CifDetSeeds.set_threshold(0.2)
    assert decoder.CifDet.get_max_detections_before_nms() == 120


@pytest.mark.parametrize('by_category', [True, False])
def test_cifdet_gpu_nms_equals_torchvision(by_category):
    """decoder/cifdet.py:55-64 on the GPU == torchvision on the host, on the same raw detections"""
    torchvision = pytest.importorskip('torchvision')
    fields = [synth.make_det_fields(6, 31, 31, n, 400 + n, n_distractors=8, n_overlapping=5)['field'] for n in (12, 60)]
    d = decoder.CifDet()
    dev = torch.from_numpy(np.stack(fields)).cuda()
    raw = d.decode_batch(dev, 16)
    got = d.decode_batch(dev, 16, nms=True, iou_threshold=0.5, nms_by_category=by_category, suppression=0.1,
                         instance_threshold=0.15)
    n_suppressed = 0
    for (cats, scores, boxes), (gc, gs, gb) in zip(raw, got):
        scores = scores.clone()
        if by_category:
            keep = torchvision.ops.batched_nms(boxes, scores, cats, 0.5)
        else:
            keep = torchvision.ops.nms(boxes, scores, 0.5)
        pre = scores.clone()
        scores *= 0.1
        scores[keep] = pre[keep]
        mask = scores > 0.15
        n_suppressed += int((~mask).sum())
        assert torch.equal(gc, cats[mask]) and torch.equal(gs, scores[mask]) and torch.equal(gb, boxes[mask])
    assert n_suppressed > 0


def test_cifdet_empty_field_and_errors():
    d = decoder.CifDet()
    cats, scores, boxes = d.call(torch.zeros((4, 6, 5, 9)), 16)
    assert cats.shape == (0,) and boxes.shape == (0, 4)
    with pytest.raises(RuntimeError):
        d.call(torch.zeros((4, 5, 5, 9)), 16)              # needs 6 components
    with pytest.raises(RuntimeError):
        d.call(torch.zeros((7, 6, 5, 9)), 16)              # category count fixed by the first call
