"""Decoder plugin for the reference framework (imported only from ``register()``, i.e. only when the
reference package ``openpifpaf`` is installed and is importing its plugins).

Reference interfaces implemented (paths relative to /root/reference/src/openpifpaf/):
  Decoder.cli / configure / factory / __call__ / batch     decoder/decoder.py:21-145
  CifCaf (the decoder this one replaces)                   decoder/cifcaf.py:81-277
  decoder selection by priority                            decoder/factory.py:110-160

Two levels of drop-in:
  * ``CifCafB200.__call__(fields)``: same contract as ``CifCaf.__call__`` (fields of one image, CPU or CUDA
    tensors) -> ``List[Annotation]``; decoding happens on the GPU through libpifpaf_b200.
  * ``CifCafB200.batch(model, image_batch, device=...)``: the whole hot loop on the GPU -- the model is
    compiled once per input shape into tcgen05 kernels and the head tensors never visit the host.
    ``register()`` makes ``Multi.batch`` (what ``Predictor`` calls, predictor.py:131) delegate to it when
    it is the selected decoder.
"""
import logging
import time

import numpy as np
import torch

from . import decoder as b200_decoder
from . import predictor as b200_predictor

LOG = logging.getLogger(__name__)


def _build_class():
    import openpifpaf
    from openpifpaf import headmeta
    from openpifpaf.annotation import Annotation
    from openpifpaf.decoder import Decoder

    class CifCafB200(Decoder):
        """CifCaf decoding on a B200 (same CLI knobs as the reference's CifCaf)."""
        fast_batch = True

        def __init__(self, cif_metas, caf_metas):
            super().__init__()
            self.cif_metas = cif_metas
            self.caf_metas = caf_metas
            self.score_weights = cif_metas[0].score_weights
            self.native = b200_decoder.CifCaf(
                len(cif_metas[0].keypoints),
                torch.LongTensor(caf_metas[0].skeleton) - 1,          # decoder/cifcaf.py:119-122
            )
            # outrank the reference's CPU CifCaf (decoder/cifcaf.py:123-125 uses the same base + n_fields/1000)
            self.priority += 1.0
            self.priority += sum(m.n_fields for m in cif_metas) / 1000.0
            self.priority += sum(m.n_fields for m in caf_metas) / 1000.0
            self._compiled = {}

        @classmethod
        def cli(cls, parser):
            group = parser.add_argument_group('CifCafB200 decoder')
            group.add_argument('--b200-no-fast-batch', dest='b200_fast_batch', default=True, action='store_false',
                               help='decode on the GPU but keep the reference model forward / host transfer')

        @classmethod
        def configure(cls, args):
            """Same argparse namespace the reference's CifCaf.configure consumes (decoder/cifcaf.py:174-211,
            decoder/factory.py:52-82); values are snapshotted into the native statics."""
            cls.fast_batch = getattr(args, 'b200_fast_batch', True)
            C = b200_decoder.CifCaf
            kp_th = getattr(args, 'keypoint_threshold', 0.15)
            kp_th_rel = getattr(args, 'keypoint_threshold_rel', 0.5)
            kp_th_nms = kp_th
            force = getattr(args, 'force_complete_pose', False)
            if force:
                if not getattr(args, 'ablation_independent_kp', False):
                    kp_th = 0.0
                kp_th_rel = 0.0
                kp_th_nms = 0.0
            seed_th = getattr(args, 'seed_threshold', 0.2)
            kp_th = min(kp_th, seed_th)
            C.set_force_complete(force)
            C.set_force_complete_caf_th(getattr(args, 'force_complete_caf_th', 0.001))
            C.set_keypoint_threshold(kp_th)
            C.set_keypoint_threshold_rel(kp_th_rel)
            C.set_greedy(getattr(args, 'greedy', False))
            C.set_block_joints(getattr(args, 'cifcaf_block_joints', False))
            b200_decoder.NMSKeypoints.set_keypoint_threshold(kp_th_nms)
            inst_th = getattr(args, 'instance_threshold', None)
            if inst_th is None:
                inst_th = 0.0 if force else 0.15
            b200_decoder.NMSKeypoints.set_instance_threshold(inst_th)
            b200_decoder.CifHr.set_threshold(getattr(args, 'cif_th', 0.3))
            b200_decoder.CifSeeds.set_threshold(seed_th)
            b200_decoder.CafScored.set_default_score_th(getattr(args, 'caf_th', 0.3))
            b200_decoder.CifSeeds.set_ablation_nms(getattr(args, 'ablation_cifseeds_nms', False))
            b200_decoder.CifSeeds.set_ablation_no_rescore(getattr(args, 'ablation_cifseeds_no_rescore', False))
            b200_decoder.CafScored.set_ablation_no_rescore(getattr(args, 'ablation_caf_no_rescore', False))

        @classmethod
        def factory(cls, head_metas):
            return [
                CifCafB200([meta], [meta_next])
                for meta, meta_next in zip(head_metas[:-1], head_metas[1:])
                if isinstance(meta, headmeta.Cif) and isinstance(meta_next, headmeta.Caf)
            ]

        def _annotations(self, ann_t, ids_t):
            """decoder/cifcaf.py:262-272"""
            out = []
            for ann_data, ann_id in zip(ann_t, ids_t):
                ann = Annotation(self.cif_metas[0].keypoints, self.caf_metas[0].skeleton,
                                 score_weights=self.score_weights)
                ann.data[:, :2] = ann_data[:, 1:3]
                ann.data[:, 2] = ann_data[:, 0]
                ann.joint_scales[:] = ann_data[:, 3]
                if ann_id != -1:
                    ann.id_ = int(ann_id)
                out.append(ann)
            return out

        def __call__(self, fields, initial_annotations=None):
            init_t = ids_t = None
            if initial_annotations:
                n = len(initial_annotations)
                init_t = torch.empty((n, self.cif_metas[0].n_fields, 4))
                ids_t = torch.empty((n,), dtype=torch.int64)
                for i, ann_py in enumerate(initial_annotations):           # decoder/cifcaf.py:228-239
                    init_t[i, :, 0] = torch.from_numpy(ann_py.data[:, 2].astype(np.float32))
                    init_t[i, :, 1] = torch.from_numpy(ann_py.data[:, 0].astype(np.float32))
                    init_t[i, :, 2] = torch.from_numpy(ann_py.data[:, 1].astype(np.float32))
                    init_t[i, :, 3] = torch.from_numpy(np.asarray(ann_py.joint_scales, dtype=np.float32))
                    ids_t[i] = getattr(ann_py, 'id_', -1)
            start = time.perf_counter()
            ann_t, ids = self.native.call_with_initial_annotations(
                fields[self.cif_metas[0].head_index], self.cif_metas[0].stride,
                fields[self.caf_metas[0].head_index], self.caf_metas[0].stride, init_t, ids_t)
            LOG.debug('b200 annotations = %d (%.1fms)', len(ann_t), (time.perf_counter() - start) * 1000.0)
            return self._annotations(ann_t.numpy(), ids.numpy())

        def batch(self, model, image_batch, *, device=None, gt_anns_batch=None):
            """decoder/decoder.py:114-137 with everything on the GPU."""
            if not self.fast_batch or not torch.cuda.is_available():
                return super().batch(model, image_batch, device=device, gt_anns_batch=gt_anns_batch)
            start = time.perf_counter()
            shell = model.module if hasattr(model, 'module') else model     # DataParallel (predictor.py:33-37)
            b, _, h, w = image_batch.shape
            key = (id(shell), h, w)
            pred = self._compiled.get(key)
            if pred is None or pred.net.max_batch < b:
                dev_index = torch.device(device).index if device is not None else 0
                pred = b200_predictor.from_shell(shell, h, w, max(b, 1), device=dev_index or 0)
                pred.decoder = self.native if self.native.device == pred.decoder.device else pred.decoder
                self._compiled[key] = pred
            results = pred.batch(image_batch if image_batch.dtype == torch.float32 else image_batch.float())
            self.last_nn_time = self.last_decoder_time = time.perf_counter() - start
            return [self._annotations(a.numpy(), i.numpy()) for a, i in results]

    return openpifpaf, CifCafB200


_CLASS = None


def decoder_class():
    global _CLASS
    if _CLASS is None:
        _CLASS = _build_class()[1]
    return _CLASS


def register():
    openpifpaf, cls = _build_class()
    global _CLASS
    _CLASS = cls
    openpifpaf.DECODERS.add(cls)

    # Predictor calls Multi.batch (predictor.py:131); let it delegate to the GPU-resident batch path when the
    # selected decoder provides one.
    Multi = openpifpaf.decoder.multi.Multi
    if not getattr(Multi, '_b200_patched', False):
        original = Multi.batch

        def batch(self, model, image_batch, *, device=None, gt_anns_batch=None):
            decs = [d for d in self.decoders if d is not None]
            if len(decs) == 1 and isinstance(decs[0], cls) and cls.fast_batch:
                res = decs[0].batch(model, image_batch, device=device, gt_anns_batch=gt_anns_batch)
                self.last_nn_time, self.last_decoder_time = decs[0].last_nn_time, 0.0
                return res
            return original(self, model, image_batch, device=device, gt_anns_batch=gt_anns_batch)

        Multi.batch = batch
        Multi._b200_patched = True
