#!/usr/bin/env python
"""bench.py -- images/sec end-to-end (backbone + heads + CifCaf decode), 641 px, batch 64 per GPU.

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --gpus N ...            # the reference's CPU path on host cores

One "step" = one pass of the hot path over one batch of synthetic input:
  images [64,3,641,641] f32 -> shufflenetv2k16 backbone + CIF/CAF heads -> batched CifCaf decode.
`value`  : images/s with the image batch already resident in HBM (CUDA events, max over ranks).
`e2e`    : the same through Predictor.batch() with HOST (pinned) images: H2D + forward + decode + D2H.
Under torchrun every rank runs an independent replica on its own GPU (images shard by rank, no
data-path collective: SURVEY.md 8e) and value is the sum over ranks / max time ("weak" scaling).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = 'shufflenetv2k16 cocokp(17kp/19caf) 641x641 bs64 per GPU'
SIZE = 641
METRIC = 'images/sec end-to-end (backbone+heads+decode) 641px bs64; decoder-only ms/img'


def peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {'hbm_gbs': p['hbm_gbs'], 'bf16_tflops': p.get('bf16_tflops_sustained', p['bf16_tflops']),
                'source': 'measured (MEASURED_PEAKS.json)'}
    return {'hbm_gbs': 6650.0, 'bf16_tflops': 1400.0, 'source': 'fallback (B200_PROFILING.md)'}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', f'--id={self.gpu}', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits', '-lms', '20'],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:       # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:       # noqa: BLE001
            self.proc.kill()
        sm, smax, reasons = [], None, set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for line in self.lines:
            parts = [p.strip() for p in line.split(',')]
            if len(parts) < 8:
                continue
            try:
                sm.append(float(parts[1]))
                smax = float(parts[2])
            except ValueError:
                continue
            for nm, v in zip(names, parts[4:8]):
                if v.lower().startswith('active'):
                    reasons.add(nm)
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': smax,
                'reasons': sorted(reasons), 'samples': len(sm)}


def dist_setup(n_gpus):
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    return rank, world, local


def barrier(world):
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    torch.cuda.synchronize()


def max_over_ranks(value, world, device):
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([value], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())
    return value


def measured_traffic(kernel, launches_per_step):
    """DRAM bytes per launch of `kernel` from the committed ncu capture of this same command (profiles/), or None.
    Only accepted if the capture holds a whole number of steps of the kernel's launches."""
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'profiles', 'r1_dram_traffic_bench_step.json')
    try:
        k = json.load(open(path))['kernels'][kernel]
    except (OSError, KeyError, ValueError):
        return None, 'no ncu capture committed'
    if k['launches'] % launches_per_step != 0:
        return None, 'capture does not cover whole steps'
    return round(k['dram_bytes_per_launch']), 'profiles/r1_dram_traffic_bench_step.json (ncu, bs64 step)'


# ----------------------------------------------------------------------------- this repo's arm
def run_b200(args):
    from openpifpaf_b200 import _lib, constants, network, predictor as pred_mod, synth
    rank, world, local = dist_setup(args.gpus)
    device = torch.device('cuda', local)
    torch.cuda.set_device(device)
    B = args.batch
    # random-init weights; heads centred and rescaled on the bench resolution so that the decoder sees the field
    # statistics of a ~5-person COCO image (isolated cells above the thresholds), not saturated noise
    plan = network.random_plan('shufflenetv2k16', seed=0, confidence_bias=-2.5)
    network.calibrate_random_heads(plan, device=local, size=SIZE, batch=2)
    net = network.CompiledNet(plan, SIZE, SIZE, B, device=local)
    predictor = pred_mod.Predictor(net, constants.COCO_N_KEYPOINTS, constants.COCO_PERSON_SKELETON, device=local)

    g = torch.Generator().manual_seed(1234 + rank)
    if args.raw_input:      # raw uint8 HWC images; ToTensor + Normalize run inside the stem kernel (not the default:
        # the reference's Predictor.batch takes the normalised float batch, and so does the headline number)
        host_images = torch.randint(0, 256, (B, SIZE, SIZE, 3), generator=g, dtype=torch.uint8).pin_memory()
    else:
        host_images = torch.randn((B, 3, SIZE, SIZE), generator=g, dtype=torch.float32).pin_memory()
    dev_images = host_images.to(device)
    stream = torch.cuda.current_stream(device)

    # ---- device-resident throughput (value)
    for _ in range(args.warmup):
        predictor.batch_device(dev_images)
    barrier(world)
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = _lib.lib().pifpaf_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        predictor.batch_device(dev_images)
    e1.record(stream)
    barrier(world)
    ms_total = e0.elapsed_time(e1)
    launches = int(_lib.lib().pifpaf_launch_count() - launches0)
    clocks = sampler.stop()
    ms_total = max_over_ranks(ms_total, world, device)
    ms_per_step = ms_total / args.steps
    value = world * B / (ms_per_step * 1e-3)

    # ---- end to end through the public API with host buffers: Predictor.batches() is the pipelined
    # iterator (H2D of batch i+1 under the compute of batch i), like the reference's Predictor.dataloader()
    host_pool = [host_images, host_images.clone().pin_memory()]
    for res in predictor.batches(host_pool[i % 2] for i in range(max(2, args.warmup))):
        pass
    barrier(world)
    t0 = time.perf_counter()
    n_ann = 0
    for res in predictor.batches(host_pool[i % 2] for i in range(args.steps)):
        n_ann = sum(len(a) for a, _ in res)
    torch.cuda.synchronize()
    t_e2e = time.perf_counter() - t0
    t_e2e = max_over_ranks(t_e2e, world, device)
    e2e_value = world * B * args.steps / t_e2e
    hdr = ((3 * B + 1) * 4 + 15) // 16 * 16
    d2h_bytes = min(hdr + 512 * 1024, hdr + B * 512 * 18 * 16)       # one fixed-size async copy per step

    out = None
    if rank == 0:
        pk = peaks()
        # ---- roofline of the dominant kernel (k_gemm_tc), measured live with CUDA events per launch
        prof_images = dev_images if not args.raw_input else torch.randn((B, 3, SIZE, SIZE), device=device)
        ms_op, kind, flops, nbytes = net.forward_timed(prof_images)
        sel = kind == 1
        gemm_ms = float(ms_op[sel].sum())
        achieved_gbs = float(nbytes[sel].sum()) / (gemm_ms * 1e-3) / 1e9
        achieved_tf = float(flops[sel].sum()) / (gemm_ms * 1e-3) / 1e12
        traffic, traffic_src = measured_traffic('k_gemm_tc', int(sel.sum()))
        roofline = {
            'kernel': 'k_gemm_tc (tcgen05 1x1-conv GEMMs, %d launches/step)' % int(sel.sum()),
            'bound': 'hbm', 'achieved': round(achieved_gbs, 1), 'peak': pk['hbm_gbs'], 'unit': 'GB/s',
            'frac': round(achieved_gbs / pk['hbm_gbs'], 4), 'traffic': traffic, 'traffic_unit': 'bytes/launch (dram read+write)',
            'traffic_source': traffic_src,
            'algorithmic_bytes_per_launch': round(float(nbytes[sel].sum()) / int(sel.sum())),
            'peak_source': pk['source'],
            'tensor_tflops': round(achieved_tf, 1), 'tensor_frac_of_measured_bf16': round(achieved_tf / pk['bf16_tflops'], 4),
            'share_of_forward': round(gemm_ms / float(ms_op.sum()), 3),
            'forward_ms': round(float(ms_op.sum()), 3),
            'by_kind_ms': {'input_conv': round(float(ms_op[kind == 0].sum()), 3), 'gemm_tc': round(gemm_ms, 3),
                           'dwconv': round(float(ms_op[kind == 2].sum()), 3)},
        }
        # ---- decoder-only on planted fields (COCO-like Poisson(4)+1 people per image)
        nb = min(B, 32)
        fields = synth.make_batch('cocokp', nb, 41, 41, None, seed=77)
        cif = torch.from_numpy(fields['cif']).to(device)
        caf = torch.from_numpy(fields['caf']).to(device)
        dec = predictor.decoder
        for _ in range(3):
            dec.decode_batch(cif, 16, caf, 16)
        torch.cuda.synchronize()
        d0, d1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        d0.record(stream)
        reps = 10
        for _ in range(reps):
            dec.decode_batch_async(cif, 16, caf, 16)
        d1.record(stream)
        torch.cuda.synchronize()
        dec_ms_per_img = d0.elapsed_time(d1) / reps / nb
        n_dec = sum(len(a) for a, _ in dec.fetch())
        cpu = cpu_baseline(args, sample_images=args.cpu_sample) if world == 1 and not args.no_cpu_baseline else None
        if args.dump_ops:
            table = []
            for i, o in enumerate(net.op_desc):
                hh, ww, _ = net.tensor_shapes[o['out']] if 'out' in o else net.tensor_shapes[o['in']]
                table.append({'op': i, 'kind': o['kind'], 'out_hw': [hh, ww], 'k_cols': o.get('k_cols'),
                              'n_out': o.get('n_out', o.get('channels')), 'stride': o.get('stride'),
                              'shuffle': o.get('shuffle_src', -1) >= 0 if 'shuffle_src' in o else None,
                              'ms': round(float(ms_op[i]), 4), 'gflops': round(float(flops[i]) / 1e9, 2),
                              'gbytes': round(float(nbytes[i]) / 1e9, 4),
                              'tflops': round(float(flops[i]) / float(ms_op[i]) / 1e9, 1),
                              'gbs': round(float(nbytes[i]) / float(ms_op[i]) / 1e6, 0)})
            with open(args.dump_ops, 'w') as f:
                json.dump({'batch': B, 'ops': table}, f, indent=1)
        out = {
            'metric': METRIC, 'value': round(value, 2), 'unit': 'images/s', 'n_gpus': world, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': round(ms_per_step, 3), 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'bf16 (f32 accumulate; decoder f32/f64)', 'data': 'synthetic',
            'config': {'workload': WORKLOAD, 'batch_per_gpu': B,
                       'input': ('raw uint8 HWC images (normalisation fused into the stem), ' if args.raw_input else 'randn images, ') +
                                'random-init weights; head pre-activations centred and rescaled to N(0,1) '
                                'with confidence bias -2.5: per image ~1000 CIF cells >= 0.3, ~3500 seed candidates, '
                                '~1700 CAF entries (the counts of a ~5-person COCO image, spatially unstructured)',
                       'decoder_input': "the network's own fields", 'parallelism': f'replica x{world}, batch sharded by rank',
                       'l2': 'inputs 315 MB/step > 126 MB L2 (no explicit flush)'},
            'impl': 'b200', 'gpu_launches': launches,
            'e2e': {'value': round(e2e_value, 2), 'unit': 'images/s',
                    'h2d_bytes_per_step': int(host_images.numel() * host_images.element_size()), 'd2h_bytes_per_step': int(d2h_bytes),
                    'api': 'openpifpaf_b200.predictor.Predictor.batches(iterable of pinned host image batches)'},
            'decoder_only': {'ms_per_img': round(dec_ms_per_img, 4), 'batch': nb, 'annotations': n_dec,
                             'fields': 'planted poses, Poisson(4)+1 people/img, 41x41 cells'},
            'roofline': roofline, 'clocks': clocks, 'cpu_baseline': cpu,
            'net_gflop_per_image': round(net.flops_per_image / 1e9, 2), 'annotations_last_step': int(n_ann),
        }
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()
    if out is not None:
        print(json.dumps(out), flush=True)


# ----------------------------------------------------------------------------- CPU arms
def cpu_path_setup():
    """The reference's CPU path: its network modules restated in PyTorch (oracle/net_oracle.py; the
    reference's own Python package cannot travel to the GPU box) + its UNMODIFIED C++ decoder compiled
    from /root/reference (oracle/_ref) when present, else the plain-C oracle port."""
    from oracle import net_oracle, cifcaf as oc
    from openpifpaf_b200 import constants
    shell = net_oracle.make_shell('shufflenetv2k16', seed=0)
    skeleton = np.asarray(constants.COCO_PERSON_SKELETON, dtype=np.int64) - 1
    if oc.ref_available():
        oc.ref_configure()
        dec_kind = 'reference'
        cls = oc.load_ref().CifCaf
        inst = cls(17, torch.from_numpy(skeleton))       # warm instance, as Predictor uses it

        def decode(cif, caf):
            return inst.call(cif, 16, caf, 16)
    else:
        dec_kind = 'port'

        def decode(cif, caf):
            return oc.decode(cif.numpy(), 16, caf.numpy(), 16, skeleton, 17)
    return shell, decode, dec_kind


def cpu_step(shell, decode, images):
    with torch.no_grad():
        cif, caf = shell(images)
    n = 0
    for b in range(images.shape[0]):
        ann = decode(cif[b].contiguous(), caf[b].contiguous())
        n += len(ann[0])
    return n


def pick_cpu_threads(shell):
    """The CPU arm gets the thread count that serves it best: PyTorch-CPU convolutions at batch 4 need not scale
    to every core of a many-core host (round 1: 0.15 images/s with all 128 threads of the GPU box, 0.66 with the 8
    threads of the build container), so a short probe (one 321x321 image per candidate) picks among all cores,
    1/2, 1/4 and 1/8 of them."""
    cores = os.cpu_count() or 1
    candidates = sorted({max(1, cores // d) for d in (1, 2, 4, 8)}, reverse=True)
    probe = torch.randn((1, 3, 321, 321), generator=torch.Generator().manual_seed(7))
    best, best_dt = cores, None
    for n in candidates:
        torch.set_num_threads(n)
        with torch.no_grad():
            shell(probe)
            t0 = time.perf_counter()
            shell(probe)
            dt = time.perf_counter() - t0
        if best_dt is None or dt < best_dt:
            best, best_dt = n, dt
    torch.set_num_threads(best)
    return best, candidates


def cpu_baseline(args, sample_images=4):
    shell, decode, dec_kind = cpu_path_setup()
    cores, _ = pick_cpu_threads(shell)
    images = torch.randn((sample_images, 3, SIZE, SIZE), generator=torch.Generator().manual_seed(1234))
    cpu_step(shell, decode, images[:1])                      # warm-up
    t0 = time.perf_counter()
    cpu_step(shell, decode, images)
    dt = time.perf_counter() - t0
    return {'value': round(sample_images / dt, 3), 'unit': 'images/s', 'cores': cores,
            'kind': 'port' if dec_kind == 'port' else 'reference',
            'sample': f'{sample_images} images 641x641: PyTorch-CPU fp32 forward of the same architecture '
                      f'(port of the reference modules, {cores} threads = the fastest of all / half / quarter / eighth of '
                      f'the {os.cpu_count()} host cores) + {dec_kind} C++ CifCaf decoder, single pass'}


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    shell, decode, dec_kind = cpu_path_setup()
    cores, _ = pick_cpu_threads(shell)
    sample = args.cpu_sample
    images = torch.randn((sample, 3, SIZE, SIZE), generator=torch.Generator().manual_seed(1234))
    for _ in range(min(args.warmup, 1)):
        cpu_step(shell, decode, images[:1])
    steps = min(args.steps, 3)
    t0 = time.perf_counter()
    for _ in range(steps):
        cpu_step(shell, decode, images)
    dt = time.perf_counter() - t0
    value = sample * steps / dt
    desc = (f'{sample} images per step, {steps} steps: PyTorch-CPU fp32 forward (port of the reference modules, '
            f'{cores} threads = the fastest of all / half / quarter / eighth of the {os.cpu_count()} host cores) + '
            f'{dec_kind} C++ CifCaf decoder (serial per image, decoder/decoder.py:33-34)')
    out = {
        'metric': METRIC, 'value': round(value, 3), 'unit': 'images/s', 'n_gpus': args.gpus, 'steps': steps,
        'warmup': min(args.warmup, 1), 'ms_per_step': round(dt / steps * 1e3, 2), 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic', 'impl': 'reference',
        'config': {'workload': WORKLOAD, 'sample': desc},
        'cpu_baseline': {'value': round(value, 3), 'unit': 'images/s', 'cores': cores,
                         'kind': 'reference' if dec_kind == 'reference' else 'port', 'sample': desc},
        'e2e': {'value': round(value, 3), 'unit': 'images/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--batch', type=int, default=64, help='images per GPU per step')
    ap.add_argument('--cpu-sample', type=int, default=4, help='images in the bounded CPU-baseline sample')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--raw-input', action='store_true',
                    help='feed raw uint8 [B,H,W,3] images (normalisation fused into the stem) instead of float32 [B,3,H,W]')
    ap.add_argument('--dump-ops', default=None, help='write the per-op timing table (profiling pass) to this JSON file')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == 'b200' else args.warmup
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_b200(args)


if __name__ == '__main__':
    main()
