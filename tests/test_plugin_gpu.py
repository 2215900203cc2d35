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
    assert sum(report.values()) > 0, report

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
