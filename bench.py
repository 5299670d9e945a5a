#!/usr/bin/env python
"""Benchmark of the hot path: ImageNet ResNet-18, ls-1 weights / ls-2 activations, forward images/s.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B]
  (N > 1: python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...)

One step = one forward of a [B,3,224,224] synthetic batch per GPU through QResNet (reference
examples/imagenet/imagenet_ls1_weight_ls2_activation_kd.yaml, random-init weights, calibrated
buffers).  Rank 0 prints ONE JSON line:
  value         images/s over all GPUs, inputs resident in HBM, whole forward as one CUDA graph,
                followed (N>1) by the NCCL all_gather of logits; CUDA events, max over ranks
  e2e           the same through ml_quant_b200.runtime.HostPipeline with HOST batches: pinned H2D of
                every batch and D2H of the logits inside the timed region (copy/compute overlapped)
  roofline      dominant kernel of the step (share from a per-launch CUDA-event pass)
  cpu_baseline  the oracle (torch-CPU restatement of the reference) on this box's host cores, bounded
--impl reference times that CPU restatement itself (the reference is pure PyTorch and cannot travel).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

CONFIG = 'imagenet_resnet18_ls1w_ls2a'
METRIC = 'resnet18_ls1w_ls2a_fwd_images_per_sec'


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get('hbm_gbs', 6650.0), d.get('bf16_tflops', 1590.0), d.get('bf16_tflops_sustained', 1400.0), 'measured'
    return 6650.0, 1590.0, 1400.0, 'fallback'


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


def cpu_reference_arm(batch, steps, warmup):
    """The reference's algorithm on the host cores (oracle/, torch CPU ops, all threads)."""
    from oracle import lsq_oracle as O
    from ml_quant_b200 import configs, runtime
    torch.set_num_threads(os.cpu_count() or 1)
    model = runtime.build_model(CONFIG)
    # buffers a checkpoint would carry: weight scales from the weights, BN stats left at init + noise
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    for k in list(sd):
        if k.endswith('w_approximate.v1'):
            w = sd[k[:-len('w_approximate.v1')] + 'weight']
            sd[k] = w.abs().mean(dim=(1, 2, 3))
    arch = configs.arch(CONFIG)
    g = torch.Generator().manual_seed(1234)
    x = torch.randn(batch, 3, 224, 224, generator=g)
    with torch.no_grad():
        for _ in range(warmup):
            O.resnet_forward(sd, arch, x)
        t = time.time()
        for _ in range(steps):
            O.resnet_forward(sd, arch, x)
        dt = time.time() - t
    return batch * steps / dt, dt / steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=50)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--batch', type=int, default=512, help='images per GPU per step')
    ap.add_argument('--no-graph', action='store_true')
    ap.add_argument('--no-fuse', action='store_true', help='keep BatchNorm / ReLU / residual adds as separate torch ops')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    hbm_peak, bf16_burst, bf16_sust, peak_src = peaks()
    base_cfg = {'workload': CONFIG + '_b%d_224' % args.batch, 'batch_per_gpu': args.batch, 'image': 224,
                'sharding': 'batch' if world > 1 else 'none',
                'l2': 'per-step input (%d MB) and every activation tensor exceed the 126 MB L2' % (args.batch * 3 * 224 * 224 * 4 >> 20)}

    if args.impl == 'reference':
        if rank != 0:
            return
        cores = os.cpu_count() or 1
        steps = max(1, min(args.steps, 3))
        sample = 64
        ips, sec = cpu_reference_arm(sample, steps, min(args.warmup, 1))
        print(json.dumps({
            'impl': 'reference', 'metric': METRIC, 'value': ips, 'unit': 'images/s', 'n_gpus': args.gpus,
            'steps': steps, 'warmup': min(args.warmup, 1), 'ms_per_step': sec * 1e3, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic', 'config': base_cfg,
            'cpu_baseline': {'value': ips, 'unit': 'images/s', 'cores': cores, 'kind': 'port',
                             'sample': f'{steps} forwards of {sample} images (oracle/lsq_oracle.py, torch CPU, {cores} threads)'},
            'e2e': {'value': ips, 'unit': 'images/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}))
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

    model = runtime.build_model(CONFIG, dev)
    runtime.calibrate(model, (3, 224, 224))
    if not args.no_fuse:
        runtime.optimize_for_inference(model)
    B = args.batch
    g = torch.Generator(device='cpu').manual_seed(1234 + rank)
    host_x = torch.randn(B, 3, 224, 224, generator=g).pin_memory()
    x = host_x.to(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    with torch.no_grad():
        ops.reset_counters()
        model(x)
        launches_per_step = sum(ops.LAUNCHES.values())
        for _ in range(max(args.warmup, 3) - 1):
            model(x)
    fwd = model if args.no_graph else runtime.GraphedForward(model, x)

    def step():
        with torch.no_grad():
            out = fwd(x)
        return runtime.gather_logits(out, world)

    for _ in range(2):
        step()
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        logits = step()
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    clocks = sampler.summary() if sampler else None
    value = world * B * args.steps / (ms / 1e3)

    # ---- end to end: host batches through the public pipeline (H2D + forward + D2H every step) ----
    pipe = runtime.HostPipeline(model, (B, 3, 224, 224), dev, use_graph=not args.no_graph)
    batches = [host_x] * args.steps
    pipe.run(batches[:2])
    barrier()
    t0 = time.perf_counter()
    pipe.run(batches)
    if world > 1:
        runtime.gather_logits(pipe.bufs[0][:1].new_zeros(B, 1000), world)
    barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = world * B * args.steps / float(e2e_s.item())

    # ---- per-launch timing of our kernels (eager pass, CUDA events on the launching stream) ----
    ops.PROFILE = []
    torch.cuda.synchronize()
    pe0, pe1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.no_grad():
        pe0.record()
        model(x)
        pe1.record()
    torch.cuda.synchronize()
    eager_ms = pe0.elapsed_time(pe1)
    agg = {}
    for name, a, b, nbytes, nops in ops.PROFILE:
        d = agg.setdefault(name, {'ms': 0.0, 'launches': 0, 'bytes': 0.0, 'ops': 0.0})
        d['ms'] += a.elapsed_time(b)
        d['launches'] += 1
        d['bytes'] += nbytes
        d['ops'] += nops
    ops.PROFILE = None
    stem_ms = None
    if hasattr(model, '_lsq_stem'):
        se0, se1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.no_grad():
            model._lsq_stem(x)
            se0.record()
            model._lsq_stem(x)
            se1.record()
        torch.cuda.synchronize()
        stem_ms = se0.elapsed_time(se1)
    ours_ms = sum(d['ms'] for d in agg.values())
    top = max(agg, key=lambda k: agg[k]['ms'])
    td = agg[top]
    traffic, traffic_instance = None, None
    tpath = os.path.join(ROOT, 'profiles', 'roofline_traffic.json')
    if os.path.exists(tpath):
        tj = json.load(open(tpath))
        traffic, traffic_instance = tj.get(top), tj.get(top + '_instance')
    if top.startswith('bconv_tc'):
        # kind::i8 issues at twice the bf16 rate: denominator = 2 x measured bf16 (sustained, MEASURED_PEAKS.json).
        # (scripts/mb/mb_umma.cu measures 4.3 POP/s for back-to-back M128 N256 K32 tcgen05.mma on this part.)
        peak = 2.0 * bf16_sust
        ach = td['ops'] / (td['ms'] / 1e3) / 1e12
        roof = {'bound': 'tensor', 'achieved': ach, 'peak': peak, 'unit': 'TOP/s', 'frac': ach / peak}
    else:
        ach = td['bytes'] / (td['ms'] / 1e3) / 1e9
        roof = {'bound': 'hbm', 'achieved': ach, 'peak': hbm_peak, 'unit': 'GB/s', 'frac': ach / hbm_peak}
    roof.update({'traffic': traffic, 'traffic_instance': traffic_instance, 'kernel': top, 'launches': td['launches'], 'avg_ms': td['ms'] / td['launches'],
                 'share_of_step': td['ms'] / (ms / args.steps), 'peak_source': peak_src,
                 'kernels_ms': {k: round(v['ms'], 3) for k, v in sorted(agg.items(), key=lambda kv: -kv[1]['ms'])},
                 'graph_step_ms': ms / args.steps, 'our_kernels_share': ours_ms / (ms / args.steps), 'fp32_stem_ms': stem_ms})
    # whole-forward HBM roofline (29.5 MB/image with ideal fusion)
    roof['forward_hbm_frac'] = (value / world) * 29.5e6 / (hbm_peak * 1e9)

    out = {
        'metric': METRIC, 'value': value, 'unit': 'images/s', 'n_gpus': world, 'steps': args.steps,
        'warmup': max(args.warmup, 3), 'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'int8', 'data': 'synthetic',
        'config': base_cfg, 'clocks': clocks,
        'e2e': {'value': e2e_value, 'unit': 'images/s', 'h2d_bytes_per_step': B * 3 * 224 * 224 * 4,
                'd2h_bytes_per_step': B * 1000 * 4},
        'gpu_launches': launches_per_step * args.steps, 'gpu_launches_per_step': launches_per_step,
        'roofline': roof, 'cuda_graph': not args.no_graph, 'fused_blocks': not args.no_fuse,
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        ips, sec = cpu_reference_arm(64, 4, 1)      # ~5-10 s of CPU work on the box's host cores
        out['cpu_baseline'] = {'value': ips, 'unit': 'images/s', 'cores': cores, 'kind': 'port',
                               'sample': f'4 forwards of 64 images (oracle/lsq_oracle.py, torch CPU, {cores} threads)'}
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
