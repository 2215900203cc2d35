"""ctypes wrapper of the plain-C oracle (oracle/cifcaf_oracle.c) and loader of the
compiled unmodified reference (oracle/_ref/refcpp.so).  TEST INFRASTRUCTURE ONLY."""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, '_build', 'libcifcaf_oracle.so')
REF_SO = os.path.join(HERE, '_ref', 'refcpp.so')


class OracleParams(ctypes.Structure):
    """Mirror of oracle_params_t."""
    _fields_ = [
        ('cifhr_neighbors', ctypes.c_int64),
        ('cifhr_threshold', ctypes.c_double),
        ('cifhr_ablation_skip', ctypes.c_int32),
        ('seed_threshold', ctypes.c_double),
        ('seeds_ablation_nms', ctypes.c_int32),
        ('seeds_ablation_no_rescore', ctypes.c_int32),
        ('caf_score_th', ctypes.c_double),
        ('caf_cif_floor', ctypes.c_double),
        ('caf_ablation_no_rescore', ctypes.c_int32),
        ('block_joints', ctypes.c_int32),
        ('greedy', ctypes.c_int32),
        ('keypoint_threshold', ctypes.c_double),
        ('keypoint_threshold_rel', ctypes.c_double),
        ('reverse_match', ctypes.c_int32),
        ('force_complete', ctypes.c_int32),
        ('force_complete_caf_th', ctypes.c_double),
        ('nms_suppression', ctypes.c_double),
        ('nms_instance_threshold', ctypes.c_double),
        ('nms_keypoint_threshold', ctypes.c_double),
        ('occ_reduction', ctypes.c_double),
        ('occ_min_scale', ctypes.c_double),
        ('cifhr_revision', ctypes.c_double),
        ('seed_sort_stable', ctypes.c_int32),
    ]


_lib = None


def build():
    subprocess.check_call(['make', '-s', '-C', HERE])
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH) or \
                os.path.getmtime(LIB_PATH) < os.path.getmtime(os.path.join(HERE, 'cifcaf_oracle.c')):
            build()
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.oracle_cifcaf_call.restype = ctypes.c_int64
        _lib.oracle_cifseeds.restype = ctypes.c_int64
        _lib.oracle_cifdet_call.restype = ctypes.c_int64
    return _lib


def default_params(**overrides):
    p = OracleParams()
    lib().oracle_default_params(ctypes.byref(p))
    for k, v in overrides.items():
        assert hasattr(p, k), k
        setattr(p, k, v)
    return p


def _ptr(a, ctype):
    return a.ctypes.data_as(ctypes.POINTER(ctype)) if a is not None else None


def decode(cif, cif_stride, caf, caf_stride, skeleton, n_keypoints=None, params=None,
           initial_annotations=None, initial_ids=None, taps=False, cap=None):
    """Oracle equivalent of CifCaf(n_keypoints, skeleton).call_with_initial_annotations(...)
    on a FRESH instance.  cif [F,5,h,w], caf [C,8,h,w] float32 numpy; skeleton [C,2] 0-based.
    Returns (annotations [N,K,4] f32, ids [N] i64[, taps dict])."""
    cif = np.ascontiguousarray(cif, dtype=np.float32)
    caf = np.ascontiguousarray(caf, dtype=np.float32)
    skeleton = np.ascontiguousarray(skeleton, dtype=np.int64)
    F, _, h, w = cif.shape
    C, _, ch, cw = caf.shape
    K = int(n_keypoints if n_keypoints is not None else F)
    p = params if params is not None else default_params()
    n_init = 0
    ia = ii = None
    if initial_annotations is not None and len(initial_annotations):
        ia = np.ascontiguousarray(initial_annotations, dtype=np.float32)
        ii = np.ascontiguousarray(initial_ids, dtype=np.int64)
        n_init = ia.shape[0]
    if cap is None:
        cap = F * h * w + n_init + 1
    out = np.zeros((cap, K, 4), dtype=np.float32)
    ids = np.zeros((cap,), dtype=np.int64)
    H, W = (h - 1) * cif_stride + 1, (w - 1) * cif_stride + 1
    t = {}
    if taps:
        t['cifhr'] = np.zeros((F, H, W), dtype=np.float32)
        t['seeds_f'] = np.zeros((F * h * w,), dtype=np.int64)
        t['seeds_vxys'] = np.zeros((F * h * w, 4), dtype=np.float32)
        t['n_seeds'] = ctypes.c_int64(0)
        t['fwd'] = np.zeros((C, ch * cw, 7), dtype=np.float32)
        t['bwd'] = np.zeros((C, ch * cw, 7), dtype=np.float32)
        t['n_fwd'] = np.zeros((C,), dtype=np.int64)
        t['n_bwd'] = np.zeros((C,), dtype=np.int64)
        t['n_pre_nms'] = ctypes.c_int64(0)
    i64, f32 = ctypes.c_int64, ctypes.c_float
    n = lib().oracle_cifcaf_call(
        _ptr(cif, f32), i64(F), i64(h), i64(w), i64(cif_stride),
        _ptr(caf, f32), i64(C), i64(ch), i64(cw), i64(caf_stride),
        _ptr(skeleton, i64), i64(K),
        _ptr(ia, f32), _ptr(ii, i64), i64(n_init),
        ctypes.byref(p),
        _ptr(out, f32), _ptr(ids, i64), i64(cap),
        _ptr(t.get('cifhr'), f32),
        _ptr(t.get('seeds_f'), i64), _ptr(t.get('seeds_vxys'), f32), i64(F * h * w),
        ctypes.byref(t['n_seeds']) if taps else None,
        _ptr(t.get('fwd'), f32), _ptr(t.get('n_fwd'), i64),
        _ptr(t.get('bwd'), f32), _ptr(t.get('n_bwd'), i64),
        ctypes.byref(t['n_pre_nms']) if taps else None)
    n = int(n)
    assert n <= cap
    if not taps:
        return out[:n].copy(), ids[:n].copy()
    ns = int(t['n_seeds'].value)
    taps_out = {
        'cifhr': t['cifhr'],
        'seeds_f': t['seeds_f'][:ns].copy(), 'seeds_vxys': t['seeds_vxys'][:ns].copy(),
        'fwd': [t['fwd'][c, :t['n_fwd'][c]].copy() for c in range(C)],
        'bwd': [t['bwd'][c, :t['n_bwd'][c]].copy() for c in range(C)],
        'n_pre_nms': int(t['n_pre_nms'].value),
    }
    return out[:n].copy(), ids[:n].copy(), taps_out


def decode_det(field, stride, params=None, max_detections_before_nms=120, taps=False):
    """Oracle equivalent of torch.classes.openpifpaf_decoder.CifDet().call(field, stride) on a FRESH instance
    (csrc/src/cifdet.cpp:24-80).  field [F,6,h,w] float32.  params.seed_threshold plays CifDetSeeds::threshold.
    Returns (categories [N] i64, scores [N] f32, boxes [N,4] f32 (x1,y1,x2,y2)[, taps dict])."""
    field = np.ascontiguousarray(field, dtype=np.float32)
    F, ncomp, h, w = field.shape
    assert ncomp == 6
    p = params if params is not None else default_params()
    cap = F * h * w + 1
    cats = np.zeros((cap,), dtype=np.int64)
    scores = np.zeros((cap,), dtype=np.float32)
    boxes = np.zeros((cap, 4), dtype=np.float32)
    H, W = (h - 1) * stride + 1, (w - 1) * stride + 1
    t = {}
    if taps:
        t['cifhr'] = np.zeros((F, H, W), dtype=np.float32)
        t['seeds_f'] = np.zeros((F * h * w,), dtype=np.int64)
        t['seeds_vxywh'] = np.zeros((F * h * w, 5), dtype=np.float32)
        t['n_seeds'] = ctypes.c_int64(0)
    i64, f32 = ctypes.c_int64, ctypes.c_float
    n = int(lib().oracle_cifdet_call(
        _ptr(field, f32), i64(F), i64(h), i64(w), i64(stride), ctypes.byref(p), i64(max_detections_before_nms),
        _ptr(cats, i64), _ptr(scores, f32), _ptr(boxes, f32), i64(cap),
        _ptr(t.get('cifhr'), f32), _ptr(t.get('seeds_f'), i64), _ptr(t.get('seeds_vxywh'), f32), i64(F * h * w),
        ctypes.byref(t['n_seeds']) if taps else None))
    out = (cats[:n].copy(), scores[:n].copy(), boxes[:n].copy())
    if not taps:
        return out
    ns = int(t['n_seeds'].value)
    return out + ({'cifhr': t['cifhr'], 'seeds_f': t['seeds_f'][:ns].copy(),
                   'seeds_vxywh': t['seeds_vxywh'][:ns].copy()},)


def grow_connection_blend(caf, x, y, s, filter_sigmas=1.0, only_max=False):
    caf = np.ascontiguousarray(caf, dtype=np.float32)
    out = (ctypes.c_double * 4)()
    lib().oracle_grow_connection_blend(
        _ptr(caf, ctypes.c_float), ctypes.c_int64(caf.shape[0]),
        ctypes.c_double(x), ctypes.c_double(y), ctypes.c_double(s),
        ctypes.c_double(filter_sigmas), ctypes.c_int(int(only_max)), out)
    return list(out)


# ---------------------------------------------------------------- compiled reference

_ref_loaded = False

# statics of the reference classes we touch, with their defaults (csrc/src/*.cpp)
REF_DEFAULTS = {
    'greedy': False, 'keypoint_threshold': 0.15, 'keypoint_threshold_rel': 0.5,
    'reverse_match': True, 'force_complete': False, 'force_complete_caf_th': 0.001,
}


def ref_available():
    return os.path.exists(REF_SO)


def load_ref():
    """Load oracle/_ref/refcpp.so (the unmodified reference decoder) into torch."""
    global _ref_loaded
    import torch
    if not _ref_loaded:
        if not os.path.exists(REF_SO):
            from . import build_ref
            build_ref.build()
        torch.ops.load_library(REF_SO)
        torch.ops.openpifpaf.set_quiet(True)
        _ref_loaded = True
    return torch.classes.openpifpaf_decoder


def ref_configure(**kw):
    """Set reference statics (module.cpp:26-32,76-117); unspecified -> defaults."""
    import torch
    dec = load_ref()
    cfg = dict(REF_DEFAULTS)
    cfg.update({k: v for k, v in kw.items() if k in REF_DEFAULTS})
    for k, v in cfg.items():
        getattr(dec.CifCaf, 'set_' + k)(v)
    u = torch.classes.openpifpaf_decoder_utils
    u.CifHr.set_threshold(kw.get('cifhr_threshold', 0.3))
    u.CifHr.set_neighbors(kw.get('cifhr_neighbors', 16))
    u.CifSeeds.set_threshold(kw.get('seed_threshold', 0.2))
    u.CafScored.set_default_score_th(kw.get('caf_score_th', 0.3))
    u.NMSKeypoints.set_instance_threshold(kw.get('nms_instance_threshold', 0.15))
    u.NMSKeypoints.set_keypoint_threshold(kw.get('nms_keypoint_threshold', 0.15))
    u.NMSKeypoints.set_suppression(kw.get('nms_suppression', 0.00001))


def ref_decode(cif, cif_stride, caf, caf_stride, skeleton, n_keypoints=None,
               initial_annotations=None, initial_ids=None, taps=False):
    """Run the unmodified reference on a FRESH CifCaf instance (revision 1.0)."""
    import torch
    dec = load_ref()
    cif_t = torch.from_numpy(np.ascontiguousarray(cif, dtype=np.float32))
    caf_t = torch.from_numpy(np.ascontiguousarray(caf, dtype=np.float32))
    sk = torch.from_numpy(np.ascontiguousarray(skeleton, dtype=np.int64))
    K = int(n_keypoints if n_keypoints is not None else cif_t.shape[0])
    inst = dec.CifCaf(K, sk)
    ia = ii = None
    if initial_annotations is not None and len(initial_annotations):
        ia = torch.from_numpy(np.ascontiguousarray(initial_annotations, dtype=np.float32))
        ii = torch.from_numpy(np.ascontiguousarray(initial_ids, dtype=np.int64))
    ann, ids = inst.call_with_initial_annotations(cif_t, cif_stride, caf_t, caf_stride, ia, ii)
    if not taps:
        return ann.numpy().copy(), ids.numpy().copy()
    u = torch.classes.openpifpaf_decoder_utils
    hr, rev = inst.get_cifhr()
    hr = hr.clone()
    seeds = u.CifSeeds(hr, rev)
    seeds.fill(cif_t, cif_stride)
    sf, sv = seeds.get()
    cs = u.CafScored(hr, rev, -1.0, 0.1)
    cs.fill(caf_t, caf_stride, sk)
    fwd, bwd = cs.get()
    t = {'cifhr': hr.numpy().copy(), 'revision': float(rev),
         'seeds_f': sf.numpy().copy(), 'seeds_vxys': sv.numpy().copy(),
         'fwd': [x.clone().numpy() for x in fwd], 'bwd': [x.clone().numpy() for x in bwd]}
    return ann.numpy().copy(), ids.numpy().copy(), t


def ref_decode_det(field, stride):
    """Run the unmodified reference CifDet (csrc/src/cifdet.cpp) on a FRESH instance.  (CifDetHr is not exported
    as a TorchScript class, module.cpp:76-102, so there are no stage taps on this side.)"""
    import torch
    dec = load_ref()
    f_t = torch.from_numpy(np.ascontiguousarray(field, dtype=np.float32))
    cats, scores, boxes = dec.CifDet().call(f_t, stride)
    return cats.numpy().copy(), scores.numpy().copy(), boxes.numpy().copy().reshape(-1, 4)
