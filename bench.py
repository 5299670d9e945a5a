#!/usr/bin/env python
"""Benchmark of the hot path (BASELINE.json): forward images/s of the binary-quantized ResNet-18 and the solver sweep.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config c4|c2|c3|c5] [--batch B]
  (N > 1: python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...)

--config (BASELINE.json `configs`):
  c4 (default, the headline)  ImageNet ResNet-18, ls-1 weights / ls-2 activations, 512 images per GPU, 224 x 224
  c2                          CIFAR-100 ResNet-18, ls-1 / ls-2, batch 256, 32 x 32
  c3                          ImageNet ResNet-18, ls-1 / ls-1 (XNOR), batch 512
  c5                          least-squares solver sweep over the 53 conv weights of a ResNet-50, k in {ls-1, ls-2, ls-T}
One step = one forward of a synthetic batch per GPU through QResNet (random-init weights, calibrated buffers) or one
solve of every tensor for the three quantizers.  Rank 0 prints ONE JSON line:
  value         units/s over all GPUs, inputs resident in HBM, the step as one CUDA graph, followed (N > 1) by the NCCL
                all_gather of logits; CUDA events, max over ranks
  e2e           the same through the public pipeline with HOST batches: pinned H2D of every batch and D2H of the result
                inside the timed region.  Image configs: the batches are uint8 pixels (what a dataset holds), the
                reference's host-side ToTensor + Normalize is applied on the device (runtime.set_pixel_input, bit-identical
                logits); e2e_fp32 is the same through the reference's own boundary (fp32 tensors normalised on the host)
  roofline      dominant kernel of the step + `kernels`: every kernel of this repository with achieved / peak / frac and
                measured DRAM bytes over algorithmic bytes (ncu, profiles/roofline_traffic.json)
  cpu_baseline  the reference on this box's host cores, bounded sample
--impl reference times the UNMODIFIED reference package (oracle/_ref/reference_full, staged by oracle/make_ref.py;
kind "reference") on the host cores -- the oracle port (kind "port") only if the staged copy is missing.
"""
import argparse
import json
import os
import re
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

CONFIGS = {
    'c4': dict(net='imagenet_resnet18_ls1w_ls2a', batch=512, image=224, metric='resnet18_ls1w_ls2a_fwd_images_per_sec',
               planes=2, mb_per_image=29.5),
    'c2': dict(net='cifar100_resnet18_ls1w_ls2a', batch=256, image=32, metric='cifar_resnet18_ls1w_ls2a_fwd_images_per_sec',
               planes=2, mb_per_image=None),
    'c3': dict(net='imagenet_resnet18_ls1w_ls1a', batch=512, image=224, metric='resnet18_ls1w_ls1a_fwd_images_per_sec',
               planes=1, mb_per_image=29.1),
    'c5': dict(net='resnet50_conv_weights', metric='ls_solver_weight_sweep_elements_per_sec'),
}
# the 53 convolution weights of torchvision's resnet50: (cout, cin, k) x count
R50 = [((64, 3, 7), 1), ((64, 64, 1), 1), ((64, 64, 3), 3), ((64, 256, 1), 2), ((128, 128, 3), 4), ((128, 256, 1), 1),
       ((128, 512, 1), 3), ((256, 64, 1), 4), ((256, 256, 3), 6), ((256, 512, 1), 1), ((256, 1024, 1), 5),
       ((512, 128, 1), 4), ((512, 256, 1), 1), ((512, 512, 3), 3), ((512, 1024, 1), 1), ((512, 2048, 1), 2),
       ((1024, 256, 1), 6), ((1024, 512, 1), 1), ((2048, 512, 1), 3), ((2048, 1024, 1), 1)]


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get('hbm_gbs', 6650.0), d.get('bf16_tflops', 1590.0), d.get('bf16_tflops_sustained', 1400.0), 'measured'
    return 6650.0, 1590.0, 1400.0, 'fallback'


def int8_peak():
    """Back-to-back tcgen05.mma kind::i8 M128 N256 K32 on all SMs (scripts/mb/mb_umma, built by __graft_entry__.build):
    the tensor pipe's own integer peak, measured in this run.  None when the binary is missing or fails."""
    exe = os.path.join(ROOT, 'scripts', 'mb', 'mb_umma')
    if not os.path.exists(exe):
        return None
    try:
        out = subprocess.run([exe], capture_output=True, text=True, timeout=60).stdout
        m = re.search(r'i8 M128 N256 aligned.*?([0-9.]+) POP/s', out)
        return float(m.group(1)) * 1e3 if m else None          # TOP/s
    except Exception:  # noqa: BLE001
        return None


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    Q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(['nvidia-smi', '-i', str(self.index), f'--query-gpu={self.Q}',
                                      '--format=csv,noheader,nounits'], capture_output=True, text=True, timeout=5).stdout
                self.rows.append([c.strip() for c in out.strip().split(',')])
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.1)

    def summary(self):
        self.stop_flag = True
        self.join(timeout=6)
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = max([int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()] or [0])
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i] == 'Active' for r in self.rows)]
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': mx or None, 'reasons': reasons,
                'samples': len(sm)}


def cpu_reference(net, images, steps, warmup):
    """The reference (or, without the staged copy, the oracle port) on the host cores, in its own process."""
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'oracle', 'ref_arm.py'), net, str(images), str(steps), str(warmup)],
                         capture_output=True, text=True, timeout=3000)
    lines = [ln for ln in out.stdout.strip().splitlines() if ln.startswith('{')]
    if not lines:
        raise RuntimeError('reference arm failed: ' + out.stderr[-2000:])
    return json.loads(lines[-1])


def r50_weights(device, seed=0):
    g = torch.Generator(device='cpu').manual_seed(seed)
    ws = []
    for (co, ci, k), cnt in R50:
        for _ in range(cnt):
            ws.append(torch.randn(co, ci * k * k, generator=g) * (2.0 / (ci * k * k)) ** 0.5)
    return [w.to(device) for w in ws] if device is not None else ws


def solver_sweep_cpu(sample_tensors, steps):
    """Oracle restatement of the three weight quantizers on a bounded sample of the sweep (host cores)."""
    from oracle import lsq_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    ws = [w for i, w in enumerate(r50_weights(None)) if i in sample_tensors]
    n = sum(w.numel() for w in ws) * 3
    t = time.perf_counter()
    for _ in range(steps):
        for w in ws:
            w4 = w.view(w.shape[0], -1, 1, 1)
            O.quant_ls1(w4)
            O.solve_v1(w, False, 3, chunk=64)
            O.solve_v1(w, True, 3, chunk=64)
    dt = time.perf_counter() - t
    return n * steps / dt, dt / steps, n


def timed_steps(step, steps, flush, barrier, dist_max):
    """K steps, CUDA events around each (L2 flushed between steps when the inputs fit in L2), summed; max over ranks."""
    barrier()
    if flush is None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            out = step()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
    else:
        evs = []
        for _ in range(steps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = step()
            e1.record()
            evs.append((e0, e1))
        barrier()
        ms = sum(a.elapsed_time(b) for a, b in evs)
    return dist_max(ms), out


PIXEL_MEAN, PIXEL_STD = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)


def kernel_table(agg, hbm_peak, i8_peak, i8_src, tf32_peak, traffic):
    rows = []
    for name, d in sorted(agg.items(), key=lambda kv: -kv[1]['ms']):
        sec = d['ms'] / 1e3
        if name.startswith('bconv_tc'):
            ach, peak, unit, bound, src = d['ops'] / sec / 1e12, i8_peak, 'TOP/s', 'tensor', i8_src
        elif name == 'stem':
            ach, peak, unit, bound, src = d['ops'] / sec / 1e12, 2.0 * tf32_peak, 'TFLOP/s', 'tensor', \
                'useful fp32 flops against the f16 pipe (MEASURED_PEAKS.json bf16_tflops); the fp32-accurate fp16 hi/lo ' \
                'split issues 5.8 f16 products per fp32 product (2 weight halves x 8 of 6 K slots x 120 of 112 columns), and ' \
                'this K = 16 no-swizzle operand layout runs at 141 clocks per instruction, 0.43 of the pipe (profiles/r2_mb_umma_f16.txt)'
        elif name == 'pwconv':
            ach, peak, unit, bound, src = d['ops'] / sec / 1e12, tf32_peak, 'TFLOP/s', 'tensor', \
                'useful fp32 flops against the tf32 pipe (MEASURED_PEAKS.json bf16_tflops / 2); the fp32-accurate split costs ' \
                '3 tf32 products per fp32 product'
        else:
            ach, peak, unit, bound, src = d['bytes'] / sec / 1e9, hbm_peak, 'GB/s', 'hbm', 'MEASURED_PEAKS.json hbm_gbs'
        t = traffic.get(name)
        rows.append({'kernel': name, 'launches': d['launches'], 'ms': round(d['ms'], 4), 'bound': bound,
                     'achieved': round(ach, 2), 'peak': round(peak, 1), 'unit': unit, 'frac': round(ach / peak, 4),
                     'peak_source': src,
                     'dram_over_algorithmic': None if not t else round(t['dram_bytes'] / t['algorithmic_bytes'], 3),
                     'traffic_instance': None if not t else t.get('instance')})
    return rows


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=50)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--config', default='c4', choices=sorted(CONFIGS))
    ap.add_argument('--batch', type=int, default=0, help='images per GPU per step (default: the config\'s)')
    ap.add_argument('--no-graph', action='store_true')
    ap.add_argument('--no-fuse', action='store_true', help='keep BatchNorm / ReLU / residual adds as separate torch ops')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    cfg = CONFIGS[args.config]
    hbm_peak, bf16_burst, bf16_sust, peak_src = peaks()
    cores = os.cpu_count() or 1
    if args.config == 'c5':
        return sweep_main(args, rank, world, local, cfg, hbm_peak, peak_src, cores)
    B = args.batch or cfg['batch']
    img = cfg['image']
    in_mb = B * 3 * img * img * 4 / 2 ** 20
    base_cfg = {'workload': '%s_b%d_%d' % (cfg['net'], B, img), 'config': args.config, 'batch_per_gpu': B, 'image': img,
                'sharding': 'batch' if world > 1 else 'none',
                'l2': ('per-step input (%d MB) and every large activation tensor exceed the 126 MB L2' % in_mb) if in_mb > 126
                else 'L2 flushed (256 MB written) between timed steps: the %.0f MB input fits in L2' % in_mb}
    ref_images = 64 if img >= 224 else 256

    if args.impl == 'reference':
        if rank != 0:
            return
        steps = max(1, args.steps)
        # a step of the CPU arm is a bounded sample of the workload: `ref_images` images of the same distribution
        per_step = ref_images if steps <= 3 else max(8, ref_images // 4)
        r = cpu_reference(cfg['net'], per_step, steps, min(args.warmup, 1))
        base_cfg['reference_step_images'] = per_step
        print(json.dumps({
            'impl': 'reference', 'metric': cfg['metric'], 'value': r['images_per_s'], 'unit': 'images/s', 'n_gpus': args.gpus,
            'steps': steps, 'warmup': min(args.warmup, 1), 'ms_per_step': r['s_per_step'] * 1e3, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic', 'config': base_cfg,
            'cpu_baseline': {'value': r['images_per_s'], 'unit': 'images/s', 'cores': r['cores'], 'kind': r['kind'],
                             'sample': f'{steps} forwards of {per_step} images each (not the {B} of the GPU step; CPU throughput '
                                       f'does not depend on the batch), {"unmodified reference package" if r["kind"] == "reference" else "oracle port"}, '
                                       f'torch CPU, {r["cores"]} threads'},
            'e2e': {'value': r['images_per_s'], 'unit': 'images/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}))
        return

    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device; the product path has no CPU fallback (use --impl reference)')
    from ml_quant_b200 import ops, runtime
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=dev)
    torch.backends.cudnn.benchmark = True
    numa = runtime.bind_host_to_gpu(local) if world > 1 else None     # before any pinned allocation (first touch)

    model = runtime.build_model(cfg['net'], dev)
    runtime.calibrate(model, (3, img, img))
    if not args.no_fuse:
        runtime.optimize_for_inference(model)
    g = torch.Generator(device='cpu').manual_seed(1234 + rank)
    host_x = torch.randn(B, 3, img, img, generator=g).pin_memory()
    x = host_x.to(dev)
    flush = None if in_mb > 126 else torch.empty(256 * 2 ** 20 // 4, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def dist_max(v):
        t = torch.tensor([v], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    with torch.no_grad():
        model(x)                                       # one-time work of a first forward (weight packing) stays out of the count
        ops.reset_counters()
        model(x)
        launches_per_step = sum(ops.LAUNCHES.values())
        for _ in range(max(args.warmup, 3) - 2):
            model(x)
    fwd = model if args.no_graph else runtime.GraphedForward(model, x)

    def step():
        with torch.no_grad():
            out = fwd(x)
        return runtime.gather_logits(out, world)

    for _ in range(2):
        step()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    ms, _ = timed_steps(step, args.steps, flush, barrier, dist_max)
    clocks = sampler.summary() if sampler else None
    value = world * B * args.steps / (ms / 1e3)

    # ---- end to end: host batches through the public pipeline (H2D + forward + D2H every step) ----
    pipe = runtime.HostPipeline(model, (B, 3, img, img), dev, use_graph=not args.no_graph)
    batches = [host_x] * args.steps
    pipe.run(batches[:2])
    barrier()
    t0 = time.perf_counter()
    pipe.run(batches)
    if world > 1:
        runtime.gather_logits(pipe.bufs[0][:1].new_zeros(B, pipe.host_out.shape[1]), world)
    barrier()
    e2e_value = world * B * args.steps / dist_max(time.perf_counter() - t0)

    # ---- the same with uint8 pixel batches (the additional entry point runtime.set_pixel_input: ToTensor + Normalize
    #      applied on the device, bit-identical to the host transform; 1 byte per pixel over PCIe instead of 4) ----
    runtime.set_pixel_input(model, PIXEL_MEAN, PIXEL_STD)
    host_u8 = torch.randint(0, 256, (B, 3, img, img), dtype=torch.uint8, generator=g).pin_memory()
    pipe8 = runtime.HostPipeline(model, (B, 3, img, img), dev, use_graph=not args.no_graph, dtype=torch.uint8)
    batches8 = [host_u8] * args.steps
    pipe8.run(batches8[:2])
    barrier()
    t0 = time.perf_counter()
    pipe8.run(batches8)
    if world > 1:
        runtime.gather_logits(pipe.bufs[0][:1].new_zeros(B, pipe8.host_out.shape[1]), world)
    barrier()
    e2e_u8_value = world * B * args.steps / dist_max(time.perf_counter() - t0)

    # ---- per-launch timing of our kernels (eager pass, CUDA events on the launching stream) ----
    ops.PROFILE = []
    torch.cuda.synchronize()
    with torch.no_grad():
        model(x)
    torch.cuda.synchronize()
    agg = {}
    for name, a, b, nbytes, nops in ops.PROFILE:
        d = agg.setdefault(name, {'ms': 0.0, 'launches': 0, 'bytes': 0.0, 'ops': 0.0})
        d['ms'] += a.elapsed_time(b)
        d['launches'] += 1
        d['bytes'] += nbytes
        d['ops'] += nops
    ops.PROFILE = None
    out = {
        'metric': cfg['metric'], 'value': value, 'unit': 'images/s', 'n_gpus': world, 'steps': args.steps,
        'warmup': max(args.warmup, 3), 'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'int8', 'data': 'synthetic',
        'config': base_cfg, 'clocks': clocks,
        # headline e2e: host batches as datasets hold them (uint8 pixels) through runtime.set_pixel_input -- ToTensor +
        # Normalize applied on the device by a 256-level table per channel, logits bit-identical to the fp32 boundary;
        # e2e_fp32: the same through the reference's own input boundary (fp32 tensors normalised on the host,
        # training.py:184-190): 4 bytes per pixel over PCIe, which caps one GPU at ~91 k images/s and 8 ranks at ~300 k
        'e2e': {'value': e2e_u8_value, 'unit': 'images/s', 'h2d_bytes_per_step': B * 3 * img * img,
                'd2h_bytes_per_step': B * pipe8.host_out.shape[1] * 4,
                'input': 'uint8 pixels [n,3,h,w] in pinned host memory; ToTensor + Normalize applied on the device through a '
                         '256-level table per channel (runtime.set_pixel_input, lsq_stem_fwd_u8 / lsq_u8_expand: additional '
                         'entry point, logits bit-identical to the fp32 boundary -- tests/test_gpu_round2.py)'},
        'e2e_fp32': {'value': e2e_value, 'unit': 'images/s', 'h2d_bytes_per_step': B * 3 * img * img * 4,
                     'd2h_bytes_per_step': B * pipe.host_out.shape[1] * 4,
                     'input': 'fp32 tensors normalised on the host (the reference\'s own input boundary, training.py:184-190)'},
        'gpu_launches': launches_per_step * args.steps, 'gpu_launches_per_step': launches_per_step,
        'cuda_graph': not args.no_graph, 'fused_blocks': not args.no_fuse, 'host_affinity': numa,
    }
    if rank == 0:
        # the tensor pipe's integer peak, measured in this run (falls back to 2 x the measured bf16 burst figure)
        i8 = int8_peak() if world == 1 else None
        i8_peak, i8_src = (i8, 'scripts/mb/mb_umma: back-to-back tcgen05.mma kind::i8 M128 N256 K32 on 148 SMs, this run') \
            if i8 else (2.0 * bf16_burst, '2 x MEASURED_PEAKS.json bf16_tflops (kind::i8 issues at twice the bf16 rate)')
        tpath = os.path.join(ROOT, 'profiles', 'roofline_traffic.json')
        traffic = json.load(open(tpath)).get('kernels', {}) if os.path.exists(tpath) else {}
        kernels = kernel_table(agg, hbm_peak, i8_peak, i8_src, bf16_burst / 2.0, traffic)
        top = kernels[0]
        ours_ms = sum(d['ms'] for d in agg.values())
        roof = {'bound': top['bound'], 'achieved': top['achieved'], 'peak': top['peak'], 'unit': top['unit'], 'frac': top['frac'],
                'traffic': (traffic.get(top['kernel']) or {}).get('dram_bytes'),
                'traffic_source': 'ncu --set full capture of one launch, kept in profiles/ (see profiles/roofline_traffic.json); '
                                  'not re-measured by this run',
                'kernel': top['kernel'], 'launches': top['launches'], 'avg_ms': top['ms'] / top['launches'],
                'share_of_step': top['ms'] / (ms / args.steps), 'peak_source': top['peak_source'] + '; hbm: ' + peak_src,
                'kernels': kernels, 'graph_step_ms': ms / args.steps,
                'our_kernels_share': ours_ms / (ms / args.steps)}
        if cfg['mb_per_image']:
            # whole-forward HBM roofline (MB per image with ideal layer fusion)
            roof['forward_hbm_frac'] = (value / world) * cfg['mb_per_image'] * 1e6 / (hbm_peak * 1e9)
            roof['forward_hbm_frac_note'] = '%.1f MB/image x images/s / measured HBM copy bandwidth' % cfg['mb_per_image']
        # bit-GEMM throughput of the binary convolutions (BASELINE.json metric, second half)
        bc = [k for k in kernels if k['kernel'].startswith('bconv_tc')]
        if bc:
            roof['bit_gemm_tops'] = bc[0]['achieved']
            roof['bit_gemm_frac_of_int8_peak'] = bc[0]['frac']
        out['roofline'] = roof
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        r = cpu_reference(cfg['net'], ref_images, 3, 1)       # ~5-15 s of CPU work on the box's host cores
        out['cpu_baseline'] = {'value': r['images_per_s'], 'unit': 'images/s', 'cores': r['cores'], 'kind': r['kind'],
                               'sample': f'3 forwards of {ref_images} images ('
                                         f'{"unmodified reference package, oracle/_ref" if r["kind"] == "reference" else "oracle port"}, '
                                         f'torch CPU, {r["cores"]} threads)'}
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def sweep_main(args, rank, world, local, cfg, hbm_peak, peak_src, cores):
    """BASELINE.json configs[4]: the three weight quantizers over the 53 conv weights of a ResNet-50 (23.45 M elements in
    26 560 rows; one multi-tensor launch per quantizer, the step captured as one CUDA graph).  N > 1: every rank solves
    its own replica of the sweep (weights are replicated in the path; weak scaling, no collective)."""
    elems = sum(co * ci * k * k * cnt for (co, ci, k), cnt in R50)
    base_cfg = {'workload': 'resnet50_conv_weights_53_tensors_ls1_ls2_lsT_skip3', 'config': 'c5', 'elements': elems,
                'rows': sum(co * cnt for (co, _, _), cnt in R50), 'sharding': 'replicas' if world > 1 else 'none',
                'l2': 'L2 flushed (256 MB written) between timed steps: the 93.8 MB of weights fit in L2'}
    unit_per_step = 3 * elems
    if args.impl == 'reference':
        if rank != 0:
            return
        sample = [0, 2, 5, 20, 27, 35, 40, 45]          # a bounded sample of the 53 tensors, every row-length class
        steps = max(1, min(args.steps, 3))
        eps, sec, n = solver_sweep_cpu(sample, steps)
        base_cfg['reference_step_elements'] = n
        print(json.dumps({
            'impl': 'reference', 'metric': cfg['metric'], 'value': eps, 'unit': 'elements/s', 'n_gpus': args.gpus,
            'steps': steps, 'warmup': 0, 'ms_per_step': sec * 1e3, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic', 'config': base_cfg,
            'cpu_baseline': {'value': eps, 'unit': 'elements/s', 'cores': cores, 'kind': 'port',
                             'sample': f'{steps} passes over {len(sample)} of the 53 tensors ({n // 3} elements) for ls-1, ls-2 and '
                                       f'ls-T (oracle/lsq_oracle.py, torch CPU, {cores} threads)'},
            'e2e': {'value': eps, 'unit': 'elements/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}))
        return
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device; the product path has no CPU fallback (use --impl reference)')
    from ml_quant_b200 import ops
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=dev)
    host_w = [w.pin_memory() for w in r50_weights(None)]
    ws = [w.to(dev) for w in host_w]
    flush = torch.empty(256 * 2 ** 20 // 4, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def dist_max(v):
        t = torch.tensor([v], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def solve_all():
        return ops.row_absmean_multi(ws), ops.solve_v1_multi(ws, False, 3), ops.solve_v1_multi(ws, True, 3)

    ops.reset_counters()
    for _ in range(max(args.warmup, 3)):
        res = solve_all()
    launches_per_step = sum(ops.LAUNCHES.values()) // max(args.warmup, 3)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream(device=dev)
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        solve_all()
    torch.cuda.current_stream().wait_stream(side)
    with torch.cuda.graph(graph):
        res = solve_all()

    def step():
        graph.replay()
        return res

    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    ms, _ = timed_steps(step, args.steps, flush, barrier, dist_max)
    clocks = sampler.summary() if sampler else None
    value = world * unit_per_step * args.steps / (ms / 1e3)
    # e2e: weights from pinned host memory, scales back to the host, every step
    host_out = [torch.empty(w.shape[0], pin_memory=True) for _ in range(3) for w in host_w]      # res = 3 groups of 53
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        for d, h in zip(ws, host_w):
            d.copy_(h, non_blocking=True)
        graph.replay()
        k = 0
        for group in res:
            for v in group:
                host_out[k].copy_(v, non_blocking=True)
                k += 1
        torch.cuda.synchronize()
    e2e_value = world * unit_per_step * args.steps / dist_max(time.perf_counter() - t0)
    # per-quantizer timing (graphs of one quantizer each would hide nothing: time the eager multi-tensor launches)
    per = {}
    for name, fn in (('ls-1', lambda: ops.row_absmean_multi(ws)), ('ls-2', lambda: ops.solve_v1_multi(ws, False, 3)),
                     ('ls-T', lambda: ops.solve_v1_multi(ws, True, 3))):
        ts = []
        for _ in range(5):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        per[name] = sorted(ts)[2]
    out = {
        'metric': cfg['metric'], 'value': value, 'unit': 'elements/s', 'n_gpus': world, 'steps': args.steps,
        'warmup': max(args.warmup, 3), 'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic', 'config': base_cfg, 'clocks': clocks,
        'e2e': {'value': e2e_value, 'unit': 'elements/s', 'h2d_bytes_per_step': elems * 4,
                'd2h_bytes_per_step': 3 * base_cfg['rows'] * 4},
        'gpu_launches': launches_per_step * args.steps, 'gpu_launches_per_step': launches_per_step, 'cuda_graph': True,
    }
    if rank == 0:
        ach = unit_per_step * 4 / (ms / args.steps / 1e3) / 1e9
        out['roofline'] = {'bound': 'hbm', 'achieved': ach, 'peak': hbm_peak, 'unit': 'GB/s', 'frac': ach / hbm_peak,
                           'traffic': None, 'kernel': 'solve_v1_multi + row_absmean_multi', 'peak_source': peak_src,
                           'note': '4 bytes per element and quantizer (one algorithmic read each), whole step',
                           'eager_ms': {k: round(v, 4) for k, v in per.items()},
                           'eager_gbs': {k: round(elems * 4 / (v / 1e3) / 1e9, 1) for k, v in per.items()}}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sample = [0, 2, 5, 20, 27, 35, 40, 45]
        eps, sec, n = solver_sweep_cpu(sample, 1)
        out['cpu_baseline'] = {'value': eps, 'unit': 'elements/s', 'cores': cores, 'kind': 'port',
                               'sample': f'1 pass over {len(sample)} of the 53 tensors ({n // 3} elements) x 3 quantizers '
                                         f'(oracle/lsq_oracle.py, torch CPU, {cores} threads)'}
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
