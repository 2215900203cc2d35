"""CPU: the plugin registers with the UNMODIFIED reference package (staged from /root/reference with the
compiled extension of oracle/_ref) and is selected by the reference's decoder factory.  Skipped when the
reference is not available (e.g. on the GPU box)."""
import os
import subprocess
import sys
import textwrap

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, 'oracle', '_ref_pkg')        # staged by oracle/build_ref.py::stage_package


def _staged():
    if not os.path.exists(os.path.join(PKG, 'openpifpaf', '_cpp.so')) and os.path.isdir('/root/reference/src/openpifpaf'):
        sys.path.insert(0, ROOT)
        from oracle import build_ref
        build_ref.stage_package()
    return os.path.exists(os.path.join(PKG, 'openpifpaf', '_cpp.so'))


@pytest.mark.skipif(not _staged(), reason='reference package not staged')
def test_plugin_registers_and_is_selected(tmp_path):
    stage = PKG
    script = textwrap.dedent('''
        import sys, warnings
        warnings.filterwarnings('ignore')
        import torch
        torch.ops.load_library = torch.ops.load_library
        import openpifpaf
        from openpifpaf.plugins.coco.constants import COCO_KEYPOINTS, COCO_PERSON_SKELETON, COCO_PERSON_SIGMAS
        assert 'openpifpaf_b200' in openpifpaf.plugin.REGISTERED, list(openpifpaf.plugin.REGISTERED)
        names = sorted(d.__name__ for d in openpifpaf.DECODERS)
        assert 'CifCafB200' in names, names
        cif = openpifpaf.headmeta.Cif('cif', 'cocokp', keypoints=COCO_KEYPOINTS, sigmas=COCO_PERSON_SIGMAS)
        caf = openpifpaf.headmeta.Caf('caf', 'cocokp', keypoints=COCO_KEYPOINTS, sigmas=COCO_PERSON_SIGMAS,
                                      skeleton=COCO_PERSON_SKELETON)
        cif.head_index, caf.head_index = 0, 1
        cif.base_stride = caf.base_stride = 16
        multi = openpifpaf.decoder.factory([cif, caf])
        top = multi.decoders[0]
        assert type(top).__name__ == 'CifCafB200', type(top).__name__
        assert top.native.n_keypoints == 17 and tuple(top.native.skeleton.shape) == (19, 2)
        # CLI flags of the reference reach the native statics
        import argparse
        parser = argparse.ArgumentParser()
        openpifpaf.decoder.cli(parser)
        args = parser.parse_args(['--force-complete-pose', '--seed-threshold=0.1'])
        openpifpaf.decoder.configure(args)
        from openpifpaf_b200 import decoder as d
        # the reference's own statics are the single source of truth; the plugin snapshots them per decode
        assert d.CifCaf.params().force_complete == 0
        type(top).sync_statics()
        p = d.CifCaf.params()
        assert (p.force_complete, p.keypoint_threshold, p.keypoint_threshold_rel, p.seed_threshold,
                p.nms_instance_threshold, p.nms_keypoint_threshold) == (1, 0.0, 0.0, 0.1, 0.0, 0.0), \\
            (p.force_complete, p.keypoint_threshold, p.seed_threshold)
        # CifDet heads select the GPU detection decoder (decoder/cifdet.py:40-47)
        det = openpifpaf.headmeta.CifDet('cifdet', 'cocodet', categories=['a', 'b', 'c'])
        det.head_index, det.base_stride = 0, 16
        args = parser.parse_args([])
        openpifpaf.decoder.configure(args)
        multi = openpifpaf.decoder.factory([det])
        assert type(multi.decoders[0]).__name__ == 'CifDetB200', type(multi.decoders[0]).__name__
        assert multi.decoders[0].native.n_categories == 3
        # --dense-connections: CifCafDense semantics (decoder/cifcaf.py:17-78), concatenated skeleton
        from openpifpaf.plugins.coco.constants import DENSER_COCO_PERSON_CONNECTIONS
        caf25 = openpifpaf.headmeta.Caf('caf25', 'cocokp', keypoints=COCO_KEYPOINTS, sigmas=COCO_PERSON_SIGMAS,
                                        skeleton=DENSER_COCO_PERSON_CONNECTIONS, sparse_skeleton=COCO_PERSON_SKELETON,
                                        only_in_field_of_view=True)
        caf25.head_index, caf25.base_stride = 2, 16
        args = parser.parse_args(['--dense-connections'])
        openpifpaf.decoder.configure(args)
        multi = openpifpaf.decoder.factory([cif, caf, caf25])
        top = multi.decoders[0]
        assert type(top).__name__ == 'CifCafDenseB200', type(top).__name__
        assert tuple(top.cifcaf.native.skeleton.shape) == (19 + len(DENSER_COCO_PERSON_CONNECTIONS), 2)
        openpifpaf.decoder.configure(parser.parse_args([]))
        # tracking heads select the tracker whose pose generator is the GPU decoder (decoder/tracking_pose.py:102-123)
        tcif = openpifpaf.headmeta.TSingleImageCif('cif', 'posetrack2018', keypoints=COCO_KEYPOINTS, sigmas=COCO_PERSON_SIGMAS)
        tcaf = openpifpaf.headmeta.TSingleImageCaf('caf', 'posetrack2018', keypoints=COCO_KEYPOINTS,
                                                   sigmas=COCO_PERSON_SIGMAS, skeleton=COCO_PERSON_SKELETON)
        ttcaf = openpifpaf.headmeta.Tcaf('tcaf', 'posetrack2018', keypoints_single_frame=COCO_KEYPOINTS,
                                         sigmas_single_frame=COCO_PERSON_SIGMAS,
                                     pose_single_frame=openpifpaf.plugins.coco.constants.COCO_UPRIGHT_POSE)
        for i, m in enumerate((tcif, tcaf, ttcaf)):
            m.head_index, m.base_stride = i, 16
        multi = openpifpaf.decoder.factory([tcif, tcaf, ttcaf])
        top = [d_ for d_ in multi.decoders if d_ is not None][0]
        assert type(top).__name__ == 'TrackingPoseB200', type(top).__name__
        gen = top.pose_generator
        assert type(gen).__name__ == 'CifCafB200' and gen.native.n_keypoints == 34
        assert tuple(gen.native.skeleton.shape) == (19 + 17, 2)
        ref_tracker = openpifpaf.decoder.TrackingPose(tcif, tcaf, ttcaf)
        assert top.priority > ref_tracker.priority
        # the synthetic tracking sequence of the GPU test does what it is meant to with the reference's own tracker:
        # every planted person becomes one track that keeps its id over the frames
        import numpy as np
        from openpifpaf_b200 import synth
        torch.ops.openpifpaf.set_quiet(True)
        frames = synth.make_tracking_sequence(33, 41, n_people=4, n_frames=4, seed=5)
        ids_by_frame = []
        for fr in frames:
            anns = ref_tracker([torch.from_numpy(fr['cif']), torch.from_numpy(fr['caf']), torch.from_numpy(fr['tcaf'])])
            assert len(anns) == 4, len(anns)
            for a in anns:
                ok = a.data[:, 2] > 0          # (posetrack2018: the tracker zeroes the ears, tracking_pose.py:42-46)
                assert ok.sum() >= 13, ok.sum()
                d2 = ((fr['keypoints'][:, ok] - a.data[None, ok, :2] / 16.0) ** 2).sum(-1).mean(-1)
                assert d2.min() < 0.05, d2.min()
            ids_by_frame.append(sorted(a.id_ for a in anns))
        assert all(ids == ids_by_frame[0] for ids in ids_by_frame), ids_by_frame
        print('PLUGIN_OK')
    ''')
    env = dict(os.environ, PYTHONPATH=f'{stage}:{ROOT}')
    r = subprocess.run([sys.executable, '-c', script], capture_output=True, text=True, env=env, cwd=str(tmp_path),
                       timeout=300)
    assert 'PLUGIN_OK' in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]
