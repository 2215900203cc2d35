"""GPU: the zero-change drop-in route on hardware.  The UNMODIFIED reference package (staged next to its compiled
extension in oracle/_ref_pkg by oracle/build_ref.py, so that it travels to the GPU box) imports this repo as the plugin
`openpifpaf_b200`; the reference's own `openpifpaf.Predictor` / `decoder.factory` / `Multi.batch` then run with
`CifCafB200` selected.  Checked against the reference's own CPU `CifCaf` (its C++ extension) decoding the SAME field
tensors: identical instance counts, xy <= 1e-4, v <= 1e-5, equal `Annotation.json_data()`.

Reference call chain exercised (paths relative to /root/reference/src/openpifpaf/):
  Predictor.__init__ / numpy_images / enumerated_dataloader   predictor.py:21-153
  decoder.factory -> Multi                                    decoder/factory.py:110-160, decoder/multi.py
  Decoder.batch (replaced), CifCaf.__call__ (the checker)     decoder/decoder.py:114-137, decoder/cifcaf.py:224-277
"""
import os
import subprocess
import sys
import textwrap

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, 'oracle', '_ref_pkg')

SCRIPT = textwrap.dedent('''
    import json, sys, warnings
    warnings.filterwarnings('ignore')
    import numpy as np
    import torch
    import openpifpaf
    torch.ops.openpifpaf.set_quiet(True)
    assert 'openpifpaf_b200' in openpifpaf.plugin.REGISTERED

    # ---- the reference's own Predictor, model from scratch (no checkpoint can be downloaded here)
    torch.manual_seed(0)
    openpifpaf.network.Factory.base_name = 'shufflenetv2k16'
    openpifpaf.network.Factory.checkpoint = None
    datamodule = openpifpaf.plugins.coco.CocoKp()
    predictor = openpifpaf.Predictor(head_metas=datamodule.head_metas)
    assert predictor.device.type == 'cuda'
    multi = predictor.processor
    top = multi.decoders[0]
    assert type(top).__name__ == 'CifCafB200', type(top).__name__
    cif_meta, caf_meta = predictor.model_cpu.head_metas
    # A from-scratch network emits saturated / empty confidence maps; give its heads trained-network-like statistics
    # (centre and rescale the head pre-activations on a probe batch, confidence bias -1) BEFORE its first use, so that
    # the force-complete pass below decodes 10-25 poses per image (checked with the reference's CPU path).
    def calibrate(model, device):
        with torch.no_grad():
            probe = torch.randn(2, 3, 161, 193, generator=torch.Generator().manual_seed(7)).to(device)
            feat = model.base_net(probe)
            mu = feat.mean((0, 2, 3))
            std = (feat - mu[None, :, None, None]).pow(2).mean().sqrt()
            for head in model.head_nets:
                conv = head.conv
                w = conv.weight / std
                conv.bias.copy_(conv.bias - w[:, :, 0, 0] @ mu)
                conv.weight.copy_(w)
                up2 = head.upsample_stride ** 2
                conv.bias.view(head.meta.n_fields, head.n_components, up2)[:, 1] += -1.0
    calibrate(predictor.model, predictor.device)
    ref_cpu = openpifpaf.decoder.CifCaf([cif_meta], [caf_meta])      # the reference's CPU decoder: the checker

    def compare(got, want, what):
        assert len(got) == len(want), (what, len(got), len(want))
        dxy = dv = 0.0
        for a, b in zip(got, want):
            assert a.data.shape == b.data.shape
            dxy = max(dxy, float(np.abs(a.data[:, :2] - b.data[:, :2]).max()))
            dv = max(dv, float(np.abs(a.data[:, 2] - b.data[:, 2]).max()))
            assert float(np.abs(np.asarray(a.joint_scales) - np.asarray(b.joint_scales)).max()) <= 1e-4
            ja, jb = a.json_data(), b.json_data()
            assert ja.keys() == jb.keys()
            assert ja['category_id'] == jb['category_id']
            assert np.allclose(ja['keypoints'], jb['keypoints'], atol=0.011), what     # json rounds to 2 decimals
            assert np.allclose(ja['bbox'], jb['bbox'], atol=0.011) and abs(ja['score'] - jb['score']) <= 0.0011
        assert dxy <= 1e-4 and dv <= 1e-5, (what, dxy, dv)
        return dxy, dv

    report = {}
    # ---- 1. Multi.batch(model, image_batch) exactly as Predictor.enumerated_dataloader calls it (predictor.py:131)
    g = torch.Generator().manual_seed(1)
    for (b, h, w) in ((3, 161, 193), (2, 257, 225), (3, 161, 193)):
        images = torch.randn(b, 3, h, w, generator=g)
        pred_batch = multi.batch(predictor.model, images, device=predictor.device)
        assert len(pred_batch) == b
        # the same fields, decoded by the reference's CPU decoder
        compiled = [p for (_, _, p) in top._compiled.values() if (p.net.in_h, p.net.in_w) == (h, w)][0]
        fields = [t.cpu() for t in compiled.net.forward(images.cuda())]
        n_total = 0
        for i in range(b):
            want = ref_cpu([f[i] for f in fields])
            dxy, dv = compare(pred_batch[i], want, f'batch {b}x{h}x{w} image {i}')
            n_total += len(want)
        report[f'{b}x{h}x{w}'] = n_total
    assert len(top._compiled) <= top.compile_cache_size          # LRU bound (two shapes seen)
    # (a from-scratch network emits no poses at the default thresholds: the counts above are usually 0 == 0; the
    # non-trivial pose comparison through the same route is part 2, where every seed becomes a completed pose)

    # ---- 2. the reference CLI statics reach the native decoder: --force-complete-pose --seed-threshold 0.1
    import argparse
    parser = argparse.ArgumentParser()
    openpifpaf.decoder.cli(parser)
    args = parser.parse_args(['--force-complete-pose', '--seed-threshold=0.1'])
    openpifpaf.decoder.configure(args)
    images = torch.randn(2, 3, 161, 193, generator=g)
    pred_batch = multi.batch(predictor.model, images, device=predictor.device)
    compiled = [p for (_, _, p) in top._compiled.values() if (p.net.in_h, p.net.in_w) == (161, 193)][0]
    fields = [t.cpu() for t in compiled.net.forward(images.cuda())]
    for i in range(2):
        want = ref_cpu([f[i] for f in fields])
        compare(pred_batch[i], want, f'force-complete image {i}')
        assert all((a.data[:, 2] > 0).all() for a in pred_batch[i])      # every pose completed
    report['force_complete'] = sum(len(p) for p in pred_batch)
    assert report['force_complete'] > 0, report
    args = parser.parse_args([])
    openpifpaf.decoder.configure(args)

    # ---- 3. the user-facing call: openpifpaf.Predictor.numpy_images (preprocess, dataloader, batch, inverse transform)
    rng = np.random.default_rng(3)
    raw = [rng.integers(0, 256, (177, 193, 3), dtype=np.uint8) for _ in range(3)]
    outs = list(predictor.numpy_images(raw))
    assert len(outs) == 3
    mean = torch.tensor([0.485, 0.456, 0.406]).view(3, 1, 1)
    std = torch.tensor([0.229, 0.224, 0.225]).view(3, 1, 1)
    for img, (pred, _, meta) in zip(raw, outs):
        x = ((torch.from_numpy(img).permute(2, 0, 1).float() / 255.0) - mean) / std     # EVAL_TRANSFORM
        direct = multi.batch(predictor.model, x.unsqueeze(0), device=predictor.device)[0]
        compare(pred, direct, 'numpy_images vs direct batch')
    report['numpy_images'] = sum(len(p) for p, _, _ in outs)
    assert predictor.total_images == 3 and predictor.last_nn_time > 0.0

    # ---- 4. single-image __call__ contract (fields on the host, like the reference's decoder workers hand them over)
    want = ref_cpu([f[0] for f in fields])
    got = top([f[0] for f in fields])
    compare(got, want, '__call__ host fields')

    # ---- 5. CifDet heads: CifDetB200 (GPU decode + GPU NMS) vs the reference's CifDet (C++ + torchvision NMS)
    from openpifpaf_b200 import synth
    det_meta = openpifpaf.headmeta.CifDet('cifdet', 'cocodet', categories=['c%d' % i for i in range(6)])
    det_meta.head_index, det_meta.base_stride = 0, 16
    det_multi = openpifpaf.decoder.factory([det_meta])
    det_top = det_multi.decoders[0]
    assert type(det_top).__name__ == 'CifDetB200', type(det_top).__name__
    ref_det = openpifpaf.decoder.CifDet([det_meta])
    n_det = 0
    for n_obj, seed in ((5, 1), (40, 2)):
        field = torch.from_numpy(synth.make_det_fields(6, 31, 35, n_obj, seed, n_distractors=6)['field'])
        want = ref_det([field.clone()])
        got = det_top([field.clone()])
        assert len(got) == len(want) and len(want) > 0, (len(got), len(want))
        for a, b in zip(got, want):
            assert a.category_id == b.category_id and abs(a.score - b.score) <= 1e-6
            assert np.abs(a.bbox - b.bbox).max() <= 1e-4
            assert a.json_data() == b.json_data()
        n_det += len(want)
    report['cifdet'] = n_det

    # ---- 6. --dense-connections: CifCafDenseB200 vs the reference's CifCafDense (decoder/cifcaf.py:17-78)
    from openpifpaf.plugins.coco.constants import (COCO_KEYPOINTS, COCO_PERSON_SKELETON, COCO_PERSON_SIGMAS,
                                                   DENSER_COCO_PERSON_CONNECTIONS)
    openpifpaf.decoder.configure(parser.parse_args(['--dense-connections']))
    cif_m = openpifpaf.headmeta.Cif('cif', 'cocokp', keypoints=COCO_KEYPOINTS, sigmas=COCO_PERSON_SIGMAS)
    caf_m = openpifpaf.headmeta.Caf('caf', 'cocokp', keypoints=COCO_KEYPOINTS, sigmas=COCO_PERSON_SIGMAS,
                                    skeleton=COCO_PERSON_SKELETON)
    caf25_m = openpifpaf.headmeta.Caf('caf25', 'cocokp', keypoints=COCO_KEYPOINTS, sigmas=COCO_PERSON_SIGMAS,
                                      skeleton=DENSER_COCO_PERSON_CONNECTIONS, sparse_skeleton=COCO_PERSON_SKELETON,
                                      only_in_field_of_view=True)
    for i, m in enumerate((cif_m, caf_m, caf25_m)):
        m.head_index, m.base_stride = i, 16
    dense_multi = openpifpaf.decoder.factory([cif_m, caf_m, caf25_m])
    dense_top = dense_multi.decoders[0]
    assert type(dense_top).__name__ == 'CifCafDenseB200', type(dense_top).__name__
    ref_dense = openpifpaf.decoder.cifcaf.CifCafDense(cif_m, caf_m, caf25_m)
    f = synth.make_fields('cocokp', 33, 41, 4, 77, n_distractors=3,
                          skeleton=list(COCO_PERSON_SKELETON) + list(DENSER_COCO_PERSON_CONNECTIONS))
    dense_fields = [torch.from_numpy(f['cif']), torch.from_numpy(f['caf'][:19].copy()),
                    torch.from_numpy(f['caf'][19:].copy())]
    want = ref_dense(dense_fields)
    got = dense_top(dense_fields)
    compare(got, want, 'dense connections')
    assert len(want) >= 4, len(want)      # (the instance threshold stays 0 after the force-complete pass: factory.py:53-57)
    report['dense'] = len(want)
    openpifpaf.decoder.configure(parser.parse_args([]))

    # ---- 7. upsample_stride 2 heads (heads.py:307-343: PixelShuffle + crop, field stride 8) through the same route
    openpifpaf.plugins.coco.CocoKp.upsample_stride = 2
    dm2 = openpifpaf.plugins.coco.CocoKp()
    torch.manual_seed(1)
    model2, _ = openpifpaf.network.Factory().factory(head_metas=dm2.head_metas)
    model2 = model2.to(predictor.device).eval()
    calibrate(model2, predictor.device)
    assert [m.stride for m in model2.head_metas] == [8, 8]
    openpifpaf.decoder.configure(parser.parse_args(['--force-complete-pose', '--seed-threshold=0.1']))
    multi2 = openpifpaf.decoder.factory(model2.head_metas)
    top2 = multi2.decoders[0]
    assert type(top2).__name__ == 'CifCafB200'
    ref2 = openpifpaf.decoder.CifCaf([model2.head_metas[0]], [model2.head_metas[1]])
    images = torch.randn(2, 3, 161, 193, generator=g)
    pred_batch = multi2.batch(model2, images, device=predictor.device)
    compiled = list(top2._compiled.values())[0][2]
    fields = [t.cpu() for t in compiled.net.forward(images.cuda())]
    assert tuple(fields[0].shape[-2:]) == (21, 25)
    with torch.no_grad():
        ref_fields = model2(images.to(predictor.device))
    assert float((ref_fields[0].cpu() - fields[0])[:, :, 1].abs().max()) < 0.08      # bf16 network vs fp32 reference
    for i in range(2):
        want = ref2([f[i] for f in fields])
        compare(pred_batch[i], want, f'upsample-2 image {i}')
    report['upsample2'] = sum(len(p) for p in pred_batch)
    assert report['upsample2'] > 0
    openpifpaf.plugins.coco.CocoKp.upsample_stride = 1
    openpifpaf.decoder.configure(parser.parse_args([]))

    # ---- 8. tracking: TrackingPoseB200 (the reference's tracker with the GPU decoder as its pose generator) against
    # the reference's TrackingPose (CPU CifCaf) over a 4-frame sequence of moving planted people
    tcif = openpifpaf.headmeta.TSingleImageCif('cif', 'posetrack2018', keypoints=COCO_KEYPOINTS, sigmas=COCO_PERSON_SIGMAS)
    tcaf = openpifpaf.headmeta.TSingleImageCaf('caf', 'posetrack2018', keypoints=COCO_KEYPOINTS, sigmas=COCO_PERSON_SIGMAS,
                                               skeleton=COCO_PERSON_SKELETON)
    ttcaf = openpifpaf.headmeta.Tcaf('tcaf', 'posetrack2018', keypoints_single_frame=COCO_KEYPOINTS,
                                     sigmas_single_frame=COCO_PERSON_SIGMAS,
                                     pose_single_frame=openpifpaf.plugins.coco.constants.COCO_UPRIGHT_POSE)
    for i, m in enumerate((tcif, tcaf, ttcaf)):
        m.head_index, m.base_stride = i, 16
    track_multi = openpifpaf.decoder.factory([tcif, tcaf, ttcaf])
    track_top = [d for d in track_multi.decoders if d is not None][0]
    assert type(track_top).__name__ == 'TrackingPoseB200', type(track_top).__name__
    assert type(track_top.pose_generator).__name__ == 'CifCafB200'
    track_ref = openpifpaf.decoder.TrackingPose(tcif, tcaf, ttcaf)
    frames = synth.make_tracking_sequence(33, 41, n_people=4, n_frames=4, seed=5)
    n_tracked = 0
    for t, fr in enumerate(frames):
        fields_t = [torch.from_numpy(fr['cif']), torch.from_numpy(fr['caf']), torch.from_numpy(fr['tcaf'])]
        want = track_ref([f.clone() for f in fields_t])
        got = track_top([f.clone() for f in fields_t])
        assert len(got) == len(want), (t, len(got), len(want))
        key = lambda a: (-a.score, a.id_)
        for a, b in zip(sorted(got, key=key), sorted(want, key=key)):
            assert np.abs(a.data - b.data).max() <= 1e-4, (t, float(np.abs(a.data - b.data).max()))
        if t >= 1:
            # identities persist: the same set of track ids relative to each tracker's first id (the id counter is global)
            ids_g = sorted(a.id_ for a in got); ids_w = sorted(a.id_ for a in want)
            assert [i - ids_g[0] for i in ids_g] == [i - ids_w[0] for i in ids_w], (ids_g, ids_w)
        n_tracked += len(want)
    assert n_tracked >= 4 * 3, n_tracked
    report['tracking'] = n_tracked
    print('PLUGIN_GPU_OK', json.dumps(report))
''')


@pytest.mark.skipif(not os.path.exists(os.path.join(PKG, 'openpifpaf', '_cpp.so')),
                    reason='oracle/_ref_pkg not staged (python oracle/build_ref.py in the build container)')
def test_reference_predictor_runs_through_the_plugin(tmp_path):
    env = dict(os.environ, PYTHONPATH=f'{PKG}:{ROOT}')
    r = subprocess.run([sys.executable, '-c', SCRIPT], capture_output=True, text=True, env=env, cwd=str(tmp_path),
                       timeout=900)
    assert 'PLUGIN_GPU_OK' in r.stdout, r.stdout[-3000:] + r.stderr[-5000:]
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    with open(os.path.join(ROOT, 'gpurun_out', 'plugin_gpu_report.txt'), 'w') as f:
        f.write(r.stdout[-2000:])
