"""GPU image preprocessing and batched annotation post-processing.

Reference (paths relative to /root/reference/src/openpifpaf/):
  Predictor._preprocess_factory        predictor.py:85-102   NormalizeAnnotations, RescaleAbsolute, CenterPad /
                                                             CenterPadTight(16), EVAL_TRANSFORM
  transforms.RescaleAbsolute / _scale  transforms/scale.py:28-100,154-176   (fast=True, OpenCV absent:
                                                             PIL.Image.resize((w, h), BILINEAR), scale.py:56-59)
  transforms.CenterPad / CenterPadTight  transforms/pad.py:15-110
  Annotation.inverse_transform / json_data / score / bbox   annotation.py:96-214

`GpuPreprocess` takes the RAW uint8 images of a batch, resizes them on the GPU bit-identically to Pillow
(pifpaf_image_resize_bilinear_u8; Pillow's ImagingResample restated, this module builds its coefficient tables) straight
into the centre of a padded uint8 canvas, and returns the canvas -- the input of `CompiledNet.forward_uint8`, whose stem
applies ToTensor + Normalize -- together with the per-image `meta` dicts the reference's transforms would have produced.
`inverse_transform_batch` / `json_data_batch` are the array forms of `Annotation.inverse_transform` / `json_data` for all
annotations of an image at once (same float32 / float64 roundings as the reference under numpy 2).
"""
import ctypes
import math

import numpy as np
import torch

from . import _lib

PRECISION_BITS = 32 - 8 - 2          # Pillow src/libImaging/Resample.c
TIGHT_PAD_FILL = (124, 116, 104)     # transforms/pad.py:100-101


def pil_bilinear_coeffs(in_size, out_size):
    """Pillow's precompute_coeffs(inSize, in0=0, in1=inSize, outSize, BILINEAR) followed by normalize_coeffs_8bpc
    (src/libImaging/Resample.c): bounds int32 [out, 2] (first input index, count) and fixed-point coefficients
    int32 [out, ksize].  IEEE double arithmetic in Pillow's order (the weight sum is accumulated left to right)."""
    in_size, out_size = int(in_size), int(out_size)
    scale = float(np.float32(in_size) - np.float32(0.0)) / out_size           # (double)(in1 - in0) / outSize, float box
    filterscale = max(scale, 1.0)
    support = 1.0 * filterscale                                               # bilinear support 1.0
    ksize = int(math.ceil(support)) * 2 + 1
    xx = np.arange(out_size, dtype=np.float64)
    center = 0.0 + (xx + 0.5) * scale
    ss = 1.0 / filterscale
    xmin = np.maximum((center - support + 0.5).astype(np.int64), 0)           # (int) truncates toward zero; clamped
    xmax = np.minimum((center + support + 0.5).astype(np.int64), in_size) - xmin
    k = np.zeros((out_size, ksize), dtype=np.float64)
    ww = np.zeros((out_size,), dtype=np.float64)
    for x in range(ksize):
        arg = ((x + xmin) - center + 0.5) * ss
        arg = np.where(arg < 0.0, -arg, arg)
        w = np.where(arg < 1.0, 1.0 - arg, 0.0)
        w = np.where(x < xmax, w, 0.0)
        k[:, x] = w
        ww = ww + w                                                           # sequential, like the C loop
    nz = ww != 0.0
    k[nz] = k[nz] / ww[nz, None]
    kk = np.where(k < 0, -0.5 + k * (1 << PRECISION_BITS), 0.5 + k * (1 << PRECISION_BITS)).astype(np.int64)
    kk = np.where(np.arange(ksize)[None, :] < xmax[:, None], kk, 0).astype(np.int32)
    bounds = np.stack([xmin, xmax], axis=1).astype(np.int32)
    return bounds, kk


def resize_bilinear_reference(image, target_w, target_h):
    """numpy restatement of Pillow's two-pass 8-bit resize with the tables above (used by the CPU parity test
    against PIL itself; the GPU kernels do the same integer arithmetic)."""
    img = np.asarray(image, dtype=np.uint8)
    h, w = img.shape[:2]

    def one_pass(a, bounds, kk, axis):
        a = np.moveaxis(a, axis, 0).astype(np.int64)
        out = np.empty((bounds.shape[0],) + a.shape[1:], dtype=np.uint8)
        for i in range(bounds.shape[0]):
            lo, n = int(bounds[i, 0]), int(bounds[i, 1])
            acc = (1 << (PRECISION_BITS - 1)) + np.tensordot(kk[i, :n].astype(np.int64), a[lo:lo + n], axes=(0, 0))
            out[i] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
        return np.moveaxis(out, 0, axis)
    if target_w != w:
        img = one_pass(img, *pil_bilinear_coeffs(w, target_w), axis=1)
    if target_h != h:
        img = one_pass(img, *pil_bilinear_coeffs(h, target_h), axis=0)
    return img


def rescale_target(w, h, long_edge):
    """RescaleAbsolute.__call__ (transforms/scale.py:163-176)"""
    s = long_edge / max(h, w)
    if h > w:
        return int(w * s), int(long_edge)
    return int(long_edge), int(h * s)


def center_pad_ltrb(w, h, target_w, target_h):
    """CenterPad.center_pad / CenterPadTight.center_pad (transforms/pad.py:37-49,87-99)"""
    left = max(0, int((target_w - w) / 2.0))
    top = max(0, int((target_h - h) / 2.0))
    right = max(0, target_w - w - left)
    bottom = max(0, target_h - h - top)
    return left, top, right, bottom


def reference_meta(w, h, target_w, target_h, ltrb):
    """The meta dict after NormalizeAnnotations -> _scale -> CenterPad (transforms/annotations.py:52-77,
    transforms/scale.py:73-96, transforms/pad.py:26-33), float64 numpy like the reference."""
    meta = {'offset': np.array((0.0, 0.0)), 'scale': np.array((1.0, 1.0)),
            'rotation': {'angle': 0.0, 'width': None, 'height': None},
            'valid_area': np.array((0.0, 0.0, w - 1, h - 1)), 'hflip': False, 'width_height': np.array((w, h))}
    if (target_w, target_h) != (w, h) or True:
        x_scale = (target_w - 1) / (w - 1)
        y_scale = (target_h - 1) / (h - 1)
        sf = np.array((x_scale, y_scale))
        meta['offset'] *= sf
        meta['scale'] *= sf
        meta['valid_area'][:2] *= sf
        meta['valid_area'][2:] *= sf
    meta['offset'] -= ltrb[:2]
    meta['valid_area'][:2] += ltrb[:2]
    return meta


class GpuPreprocess:
    """:param long_edge: RescaleAbsolute target (None: no rescale, like --long-edge unset)
    :param batched: True -> CenterPad(long_edge) squares (predictor.py:91-93, batch size > 1);
                    False -> CenterPadTight(16) (predictor.py:94-95)"""

    def __init__(self, long_edge=None, batched=True, *, device=0, multiple=16):
        if batched and not long_edge:
            raise RuntimeError('--long-edge must be provided for batch size > 1')      # predictor.py:92
        self.long_edge = int(long_edge) if long_edge else None
        self.batched = bool(batched)
        self.multiple = int(multiple)
        self.device = torch.device('cuda', device)
        self._coeff_cache = {}

    def _coeffs(self, n_in, n_out):
        key = (n_in, n_out)
        if key not in self._coeff_cache:
            b, k = pil_bilinear_coeffs(n_in, n_out)
            self._coeff_cache[key] = (torch.from_numpy(b).to(self.device), torch.from_numpy(k).to(self.device), k.shape[1])
        return self._coeff_cache[key]

    def plan(self, sizes):
        """sizes: [(w, h)] of the raw images -> per-image (target_w, target_h, ltrb) and the canvas (W, H)"""
        out = []
        for (w, h) in sizes:
            tw, th = rescale_target(w, h, self.long_edge) if self.long_edge else (w, h)
            if self.batched:
                cw = ch = self.long_edge
            else:
                cw = math.ceil((tw - 1) / self.multiple) * self.multiple + 1
                ch = math.ceil((th - 1) / self.multiple) * self.multiple + 1
            out.append((tw, th, center_pad_ltrb(tw, th, cw, ch), (cw, ch)))
        canvases = {o[3] for o in out}
        if len(canvases) != 1:
            raise RuntimeError('images of one batch must share the padded size (use batched=True with a long edge)')
        return out, canvases.pop()

    def __call__(self, images, fill=None, stream=None):
        """images: list of uint8 [h, w, 3] arrays / tensors (host).  fill: pad colour -- an int (grey) or an (r, g, b)
        triple, or a list of those per image.  Default like the reference: CenterPad draws a random grey per image
        with torch.randint(0, 255) (transforms/pad.py:52), CenterPadTight uses (124, 116, 104) (pad.py:100-101).
        Returns (canvas uint8 CUDA [B, H, W, 3], metas)."""
        imgs = [torch.as_tensor(np.ascontiguousarray(im)) if not isinstance(im, torch.Tensor) else im.contiguous()
                for im in images]
        for im in imgs:
            if im.dtype != torch.uint8 or im.dim() != 3 or im.shape[2] != 3:
                raise RuntimeError('images must be uint8 [h, w, 3]')
        sizes = [(int(im.shape[1]), int(im.shape[0])) for im in imgs]
        plans, (cw, ch) = self.plan(sizes)
        B = len(imgs)
        if fill is None:
            fill = [int(torch.randint(0, 255, (1,)).item()) if self.batched else TIGHT_PAD_FILL for _ in range(B)]
        elif isinstance(fill, int) or (isinstance(fill, tuple) and len(fill) == 3 and B != 3):
            fill = [fill] * B
        fill = [(f, f, f) if isinstance(f, int) else tuple(int(c) for c in f) for f in fill]
        st = stream if stream is not None else torch.cuda.current_stream(self.device)
        L = _lib.lib()
        sp = ctypes.c_void_p(st.cuda_stream)
        with torch.cuda.stream(st):
            canvas = torch.empty((B, ch, cw, 3), dtype=torch.uint8, device=self.device)
            metas, keep = [], []
            for b, (im, (w, h), (tw, th, ltrb, _)) in enumerate(zip(imgs, sizes, plans)):
                _lib.check(L.pifpaf_image_fill_rgb(canvas[b].data_ptr(), ch * cw, fill[b][0], fill[b][1], fill[b][2], sp))
                src = im.pin_memory().to(self.device, non_blocking=True) if not im.is_cuda else im
                xb = xk = yb = yk = None
                xks = yks = 0
                if tw != w:
                    xb, xk, xks = self._coeffs(w, tw)
                if th != h:
                    yb, yk, yks = self._coeffs(h, th)
                tmp = torch.empty((h * tw * 3,), dtype=torch.uint8, device=self.device)
                dst = canvas[b].data_ptr() + (ltrb[1] * cw + ltrb[0]) * 3
                _lib.check(L.pifpaf_image_resize_bilinear_u8(
                    src.data_ptr(), h, w, dst, cw * 3, th, tw,
                    xb.data_ptr() if xb is not None else None, xk.data_ptr() if xk is not None else None, xks,
                    yb.data_ptr() if yb is not None else None, yk.data_ptr() if yk is not None else None, yks,
                    tmp.data_ptr(), sp))
                keep.append((src, tmp))
                metas.append(reference_meta(w, h, tw, th, np.asarray(ltrb)))
            self._keepalive = keep
        return canvas, metas


# ----------------------------------------------------------------------------- annotations, all of an image at once

def inverse_transform_batch(ann, meta):
    """Annotation.inverse_transform (annotation.py:162-214; no rotation) for the packed decoder output of one image.
    ann: [N, K, 4] float32 (v, x, y, s) as the decoder returns it.  Returns (data [N, K, 3] float32 (x, y, v),
    joint_scales [N, K] float32) in original-image coordinates.  Rounding like the reference under numpy 2: the
    float32 arrays meet float64 meta scalars, each statement computes in float64 and stores float32."""
    ann = np.asarray(ann, dtype=np.float32)
    data = np.empty(ann.shape[:2] + (3,), dtype=np.float32)
    data[..., 0] = ann[..., 1]
    data[..., 1] = ann[..., 2]
    data[..., 2] = ann[..., 0]
    scales = ann[..., 3].copy()
    if meta['rotation']['angle'] != 0.0:
        raise RuntimeError('rotated inputs are not part of the inference path')
    ox, oy = np.float64(meta['offset'][0]), np.float64(meta['offset'][1])
    sx, sy = np.float64(meta['scale'][0]), np.float64(meta['scale'][1])
    data[..., 0] = data[..., 0].astype(np.float64) + ox
    data[..., 1] = data[..., 1].astype(np.float64) + oy
    data[..., 0] = data[..., 0].astype(np.float64) / sx
    data[..., 1] = data[..., 1].astype(np.float64) / sy
    scales = (scales.astype(np.float64) / sx).astype(np.float32)
    if meta['hflip']:
        w = meta['width_height'][0]
        data[..., 0] = -data[..., 0] + (w - 1)
        if meta.get('horizontal_swap'):
            for i in range(data.shape[0]):
                data[i] = meta['horizontal_swap'](data[i])
    return data, scales


def json_data_batch(data, joint_scales, score_weights=None, category_id=1, coordinate_digits=2):
    """Annotation.json_data (annotation.py:121-143) with score (annotation.py:96-111) and bbox_from_keypoints
    (annotation.py:150-160) for all annotations of an image; returns a list of dicts equal to the reference's."""
    data = np.asarray(data, dtype=np.float32)
    n, K = data.shape[:2]
    sw = np.ones((K,)) if score_weights is None else np.asarray(score_weights, dtype=np.float64).copy()
    sw = sw / np.sum(sw)
    out = []
    v_mask = data[..., 2] > 0.0
    kp = data.copy()
    kp[..., 2] = np.where(v_mask, np.maximum(np.float32(0.01), kp[..., 2]), kp[..., 2])
    kp = np.around(kp.astype(np.float64), coordinate_digits)
    js = np.asarray(joint_scales, dtype=np.float32)
    for i in range(n):
        m = data[i, :, 2] > 0
        if not np.any(m):
            bbox = [0, 0, 0, 0]
        else:
            x = np.min(data[i, :, 0][m] - js[i][m])
            y = np.min(data[i, :, 1][m] - js[i][m])
            bbox = [x, y, np.max(data[i, :, 0][m] + js[i][m]) - x, np.max(data[i, :, 1][m] + js[i][m]) - y]
        score = np.sum(sw * np.sort(data[i, :, 2])[::-1])
        out.append({'keypoints': kp[i].reshape(-1).tolist(),
                    'bbox': [round(float(c), coordinate_digits) for c in bbox],
                    'score': max(0.001, round(score, 3)),
                    'category_id': category_id})
    return out
