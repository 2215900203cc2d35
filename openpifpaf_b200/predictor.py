import os
"""End-to-end batch API: images -> backbone+heads -> CifCaf decode -> annotations.

Mirror of the reference's hot loop (paths relative to /root/reference/src/openpifpaf/):
  Predictor.enumerated_dataloader   predictor.py:118-153
  Decoder.batch / fields_batch      decoder/decoder.py:76-137

Differences by design: the head tensors never leave the GPU (the reference does
``heads.cpu()`` at decoder/decoder.py:98 and decodes on host cores); the whole
batch is decoded by one batched launch sequence; only the final annotations
(a few KB) are copied back.  With ``overlap_decode`` the decode of batch i runs
on its own stream (on SMs the forward's persistent grids leave free) under the
forward of batch i+1; the head outputs are double buffered for that.
"""
import time

import torch

from . import decoder as _decoder
from . import network as _network
from . import preprocess as _preprocess


class Predictor:
    """:param net: `network.CompiledNet`
    :param n_keypoints: number of keypoints of the CIF head
    :param skeleton: 1-based skeleton of the CAF head (``headmeta.Caf.skeleton``)
    :param cif_head, caf_head: indices of the two heads in the model output (``headmeta.head_index``)
    :param overlap_decode: decode on a second stream, concurrently with the next forward
    :param reserve_sms: SMs kept free of the forward's persistent grids for the concurrent decode
                        (None: one per image of the batch, at most 1/8 of the GPU)"""

    def __init__(self, net, n_keypoints, skeleton, *, device=0, cif_head=0, caf_head=1,
                 overlap_decode=False, reserve_sms=None):
        self.net = net
        self.device = torch.device('cuda', device)
        sk = torch.as_tensor(skeleton, dtype=torch.int64).reshape(-1, 2) - 1     # decoder/cifcaf.py:121
        self.decoder = _decoder.CifCaf(n_keypoints, sk, device=device)
        # size the native handle once, from the net (never re-created while results are in flight)
        hd = net.heads[cif_head]
        self.decoder.reserve(net.max_batch, hd['h'], hd['w'], max(net.heads[cif_head]['stride'], net.heads[caf_head]['stride']))
        self.cif_head, self.caf_head = int(cif_head), int(caf_head)
        self.stream = torch.cuda.Stream(device=self.device)
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self.dec_stream = torch.cuda.Stream(device=self.device)
        self.overlap_decode = bool(overlap_decode)
        self._dec_done = [None, None]
        self._step = 0
        if self.overlap_decode:
            net.set_head_buffers(2)
            n_sm = torch.cuda.get_device_properties(self.device).multi_processor_count
            if reserve_sms is None:
                env = os.environ.get('PIFPAF_RESERVE_SMS')
                # measured (profiles/r2_history.md, sessions r / s): at 64 images per step the network wants every SM
                # (4357 -> 4579 images/s from 18 -> 0 reserved), at 8-16 images the decode wants 8 of its own
                reserve_sms = int(env) if env is not None else (8 if net.max_batch <= 16 else 4 if net.max_batch <= 32 else 0)
            net.set_sm_limit(n_sm - int(reserve_sms) if reserve_sms else 0)
        self._dev_images = None
        self.last_nn_time = 0.0
        self.last_decoder_time = 0.0
        self.image_mean = _network.CompiledNet.IMAGE_MEAN
        self.image_std = _network.CompiledNet.IMAGE_STD
        #: benchmark hook: (cif [B,F,5,h,w], cif_stride, caf [B,C,8,h,w], caf_stride) CUDA tensors decoded INSTEAD of
        #: the network's own head outputs (random-init weights emit no poses; planted fields give the decoder the
        #: work of real images).  None in production.
        self.decode_fields_override = None

    def close(self):
        """Free the GPU buffers of the compiled net and the decoder workspace now (not at garbage collection)."""
        torch.cuda.synchronize(self.device)
        self.decoder._free()
        self.net.close()

    def fields_batch(self, image_batch):
        """decoder/decoder.py:76-112 without the .cpu(): returns device-resident head tensors."""
        return self._forward(image_batch)

    def _forward(self, image_batch_dev):
        """float32 [B,3,H,W] (already normalised, what the reference's dataloader yields) or raw uint8 [B,H,W,3]
        (ToTensor + Normalize of transforms/__init__.py:26-33 fused into the stem kernel)."""
        if image_batch_dev.dtype == torch.uint8:
            return self.net.forward_uint8(image_batch_dev, mean=self.image_mean, std=self.image_std)
        return self.net.forward(image_batch_dev)

    def _decode_inputs(self, heads, batch):
        if self.decode_fields_override is not None:
            cif, cs, caf, fs = self.decode_fields_override
            return cif[:batch], cs, caf[:batch], fs
        return (heads[self.cif_head], self.net.heads[self.cif_head]['stride'],
                heads[self.caf_head], self.net.heads[self.caf_head]['stride'])

    def batch_device(self, image_batch_dev):
        """Device-resident images -> enqueue forward + decode (no host sync).  The forward goes to the current
        stream; the decode to the current stream too, or, with overlap_decode, to the decode stream behind an event
        (call join() before timing or reusing the current stream's results)."""
        cur = torch.cuda.current_stream(self.device)
        if not self.overlap_decode:
            heads = self._forward(image_batch_dev)
            cif, cs, caf, fs = self._decode_inputs(heads, int(image_batch_dev.shape[0]))
            self.decoder.decode_batch_async(cif, cs, caf, fs)
            return
        slot = self._step & 1
        self._step += 1
        if self._dec_done[slot] is not None:
            cur.wait_event(self._dec_done[slot])       # the head-output set this forward overwrites has been decoded
        heads = self._forward(image_batch_dev)
        ready = torch.cuda.Event()
        ready.record(cur)
        self.dec_stream.wait_event(ready)
        cif, cs, caf, fs = self._decode_inputs(heads, int(image_batch_dev.shape[0]))
        self.decoder.decode_batch_async(cif, cs, caf, fs, stream=self.dec_stream)
        done = torch.cuda.Event()
        done.record(self.dec_stream)
        self._dec_done[slot] = done

    def result_stream(self):
        """The stream the decode results are produced on."""
        return self.dec_stream if self.overlap_decode else torch.cuda.current_stream(self.device)

    def join(self):
        """Make the current stream wait for every decode enqueued so far (overlap_decode)."""
        cur = torch.cuda.current_stream(self.device)
        for ev in self._dec_done:
            if ev is not None:
                cur.wait_event(ev)

    def batch(self, image_batch_host):
        """decoder/decoder.py:114-137: image batch (host, ideally pinned; float32 [B,3,H,W] normalised, or raw
        uint8 [B,H,W,3]) -> per-image (annotations [N,K,4], ids [N]) CPU tensors.  H2D, forward, decode and D2H
        all inside."""
        t0 = time.perf_counter()
        with torch.cuda.stream(self.stream):
            if (self._dev_images is None or self._dev_images.shape != image_batch_host.shape
                    or self._dev_images.dtype != image_batch_host.dtype):
                self._dev_images = torch.empty(image_batch_host.shape, dtype=image_batch_host.dtype, device=self.device)
            self._dev_images.copy_(image_batch_host, non_blocking=True)
            self.batch_device(self._dev_images)
            result = self.decoder.fetch(stream=self.result_stream())
            self.join()
        self.last_nn_time = self.last_decoder_time = time.perf_counter() - t0
        return result

    def raw_images(self, images, *, json_data=False, fill=None, score_weights=None):
        """The reference's Predictor.numpy_images / images (predictor.py:155-196) for a list of RAW uint8 [h, w, 3]
        images with everything after the JPEG decode on the GPU: RescaleAbsolute(long edge = the compiled input size)
        + CenterPad (`preprocess.GpuPreprocess`, bit-identical to the reference's Pillow path), ToTensor + Normalize in
        the stem, forward, decode; then `Annotation.inverse_transform` (and `json_data`) for all annotations of an
        image at once.  Returns per image (data [N, K, 3] (x, y, v) in original-image pixels, joint_scales [N, K], meta)
        or, with json_data, (list of dicts, meta)."""
        if self.net.in_h != self.net.in_w:
            raise RuntimeError('raw_images needs a net compiled for a square input (CenterPad(long_edge))')
        if getattr(self, '_gpu_preprocess', None) is None:
            self._gpu_preprocess = _preprocess.GpuPreprocess(self.net.in_w, batched=True, device=self.device.index)
        out = []
        for i in range(0, len(images), self.net.max_batch):
            chunk = images[i:i + self.net.max_batch]
            with torch.cuda.stream(self.stream):
                canvas, metas = self._gpu_preprocess(chunk, fill=fill, stream=self.stream)
                self.batch_device(canvas)
                result = self.decoder.fetch(stream=self.result_stream())
                self.join()
            for (ann, _), meta in zip(result, metas):
                data, scales = _preprocess.inverse_transform_batch(ann.numpy(), meta)
                if json_data:
                    out.append((_preprocess.json_data_batch(data, scales, score_weights=score_weights), meta))
                else:
                    out.append((data, scales, meta))
        return out

    def batches(self, host_batches):
        """Pipelined variant of `batch` over an iterable of host image batches (what Predictor.dataloader /
        enumerated_dataloader does in the reference, predictor.py:118-153): the H2D copy of batch i+1 runs on a
        copy stream under the forward+decode of batch i, and the (tiny) result D2H of batch i is waited for
        only after batch i+1 has been enqueued.  Yields per-batch results in order."""
        dev, copied, consumed = [None, None], [None, None], [None, None]
        outstanding = 0
        for i, host in enumerate(host_batches):
            s = i % 2
            if dev[s] is None or dev[s].shape != host.shape or dev[s].dtype != host.dtype:
                dev[s] = torch.empty(host.shape, dtype=host.dtype, device=self.device)
                copied[s], consumed[s] = torch.cuda.Event(), None
            with torch.cuda.stream(self.copy_stream):
                if consumed[s] is not None:
                    self.copy_stream.wait_event(consumed[s])
                dev[s].copy_(host, non_blocking=True)
                copied[s].record(self.copy_stream)
            with torch.cuda.stream(self.stream):
                self.stream.wait_event(copied[s])
                self.batch_device(dev[s])
                consumed[s] = torch.cuda.Event()
                consumed[s].record(self.stream)        # the forward (the only reader of dev[s]) is on self.stream
                self.decoder.fetch_begin(stream=self.result_stream())
            outstanding += 1
            if outstanding == 2:
                yield self.decoder.fetch_end()
                outstanding -= 1
        while outstanding:
            yield self.decoder.fetch_end()
            outstanding -= 1
        with torch.cuda.stream(self.stream):
            self.join()


def from_shell(shell, in_h, in_w, max_batch, *, device=0, cif_meta=None, caf_meta=None, **kwargs):
    """Compile a reference-style Shell (network/nets.py:7-48) with a (Cif, Caf) head pair into a Predictor.
    cif_meta / caf_meta: the head metas the decoder was built for (decoder/cifcaf.py:212-222); default: the
    first Cif-like meta followed by a Caf-like one."""
    plan = _network.plan_from_shell(shell)
    net = _network.CompiledNet(plan, in_h, in_w, max_batch, device=device)
    metas = list(shell.head_metas)
    if cif_meta is None or caf_meta is None:
        cif_meta, caf_meta = metas[0], metas[1]
    skeleton = getattr(caf_meta, 'skeleton', None)
    if skeleton is None:
        raise RuntimeError('CAF head meta has no skeleton')

    def index_of(meta, default):
        hi = getattr(meta, 'head_index', None)
        return int(hi) if hi is not None else default

    return Predictor(net, cif_meta.n_fields, skeleton, device=device,
                     cif_head=index_of(cif_meta, 0), caf_head=index_of(caf_meta, 1), **kwargs)
