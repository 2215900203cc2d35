"""End-to-end batch API: images -> backbone+heads -> CifCaf decode -> annotations.

Mirror of the reference's hot loop (paths relative to /root/reference/src/openpifpaf/):
  Predictor.enumerated_dataloader   predictor.py:118-153
  Decoder.batch / fields_batch      decoder/decoder.py:76-137

Differences by design: the head tensors never leave the GPU (the reference does
``heads.cpu()`` at decoder/decoder.py:98 and decodes on host cores); the whole
batch is decoded by one batched launch sequence; only the final annotations
(a few KB) are copied back.
"""
import time

import torch

from . import decoder as _decoder
from . import network as _network


class Predictor:
    """:param net: `network.CompiledNet`
    :param skeleton: 1-based skeleton of the CAF head (``headmeta.Caf.skeleton``)
    :param n_keypoints: number of keypoints of the CIF head"""

    def __init__(self, net, n_keypoints, skeleton, *, device=0):
        self.net = net
        self.device = torch.device('cuda', device)
        sk = torch.as_tensor(skeleton, dtype=torch.int64).reshape(-1, 2) - 1     # decoder/cifcaf.py:121
        self.decoder = _decoder.CifCaf(n_keypoints, sk, device=device)
        self.cif_head, self.caf_head = 0, 1
        self.stream = torch.cuda.Stream(device=self.device)
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self._dev_images = None
        self.last_nn_time = 0.0
        self.last_decoder_time = 0.0
        self.image_mean = _network.CompiledNet.IMAGE_MEAN
        self.image_std = _network.CompiledNet.IMAGE_STD

    def fields_batch(self, image_batch):
        """decoder/decoder.py:76-112 without the .cpu(): returns device-resident head tensors."""
        return self._forward(image_batch)

    def _forward(self, image_batch_dev):
        """float32 [B,3,H,W] (already normalised, what the reference's dataloader yields) or raw uint8 [B,H,W,3]
        (ToTensor + Normalize of transforms/__init__.py:26-33 fused into the stem kernel)."""
        if image_batch_dev.dtype == torch.uint8:
            return self.net.forward_uint8(image_batch_dev, mean=self.image_mean, std=self.image_std)
        return self.net.forward(image_batch_dev)

    def batch_device(self, image_batch_dev):
        """Device-resident images -> enqueue forward + decode on the current stream (no host sync)."""
        heads = self._forward(image_batch_dev)
        cif, caf = heads[self.cif_head], heads[self.caf_head]
        self.decoder.decode_batch_async(cif, self.net.heads[self.cif_head]['stride'],
                                        caf, self.net.heads[self.caf_head]['stride'])

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
            result = self.decoder.fetch(stream=self.stream)
        self.last_nn_time = self.last_decoder_time = time.perf_counter() - t0
        return result

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
                consumed[s].record(self.stream)
                self.decoder.fetch_begin(stream=self.stream)
            outstanding += 1
            if outstanding == 2:
                yield self.decoder.fetch_end()
                outstanding -= 1
        while outstanding:
            yield self.decoder.fetch_end()
            outstanding -= 1


def from_shell(shell, in_h, in_w, max_batch, *, device=0):
    """Compile a reference-style Shell (network/nets.py:7-48) whose heads are (Cif, Caf) into a Predictor."""
    plan = _network.plan_from_shell(shell)
    net = _network.CompiledNet(plan, in_h, in_w, max_batch, device=device)
    metas = shell.head_metas
    cif_meta, caf_meta = metas[0], metas[1]
    skeleton = getattr(caf_meta, 'skeleton', None)
    if skeleton is None:
        raise RuntimeError('CAF head meta has no skeleton')
    return Predictor(net, cif_meta.n_fields, skeleton, device=device)
