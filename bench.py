#!/usr/bin/env python
"""bench.py — images/sec of the SimpleAICV classification training step on B200.

Workload (BASELINE.json configs[1]): ResNet-50, 224x224, batch 256 per GPU, synthetic images,
CELoss, SGD(lr 0.1, momentum 0.9, wd 1e-4, 1-D params undecayed) — one "step" is forward + loss +
backward (+ bucketed gradient all-reduce when N > 1) + optimizer step, exactly the work of
tools/scripts.py:141-270 in the reference.

  python bench.py [--gpus N] [--steps K] [--warmup W]            our arm (sm_100a kernels)
  python bench.py --impl reference ...                           reference arm: the CPU oracle
                                                                 (oracle/, restating the reference)
Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for what every key means.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

METRICS = {'resnet50': 'images/sec (ResNet-50 224x224 training step, whole job)',
           'vit_base_patch16': 'images/sec (ViT-B/16 224x224 training step, whole job)'}
FWD_FLOPS = {'resnet50': 8.178e9, 'vit_base_patch16': 35.13e9}   # per image forward; a step is 3x
ALGO_BYTES = {'resnet50': 130e6, 'vit_base_patch16': 3 * 65e6}  # per image per step, activations once each way (8d)


def _peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return {'hbm_gbs': d['hbm_gbs'], 'tf_burst': d['bf16_tflops'], 'tf_sustained': d.get('bf16_tflops_sustained', d['bf16_tflops']),
                'source': 'measured'}
    return {'hbm_gbs': 6650.0, 'tf_burst': 1590.0, 'tf_sustained': 1400.0, 'source': 'fallback'}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.path = tempfile.mktemp(suffix='.csv')
        self.proc = None
        self.idx = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.idx), f'--query-gpu={self.Q}', '--format=csv,noheader,nounits',
                                          '-lms', '200'], stdout=open(self.path, 'w'), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], 0, set()
        for line in open(self.path):
            f = [t.strip() for t in line.split(',')]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx = max(mx, float(f[2]))
            except ValueError:
                continue
            for name, val in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), f[5:9]):
                if val.lower().startswith('active'):
                    reasons.add(name)
        sm.sort()
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': mx or None, 'reasons': sorted(reasons),
                'samples': len(sm)}


def build_optimizer(model, lr=0.1, momentum=0.9, weight_decay=1e-4):
    """tools/utils.py:292-600 with global_weight_decay=False (imagenet/resnet50/train_config.py:66-79)."""
    decay = [p for p in model.parameters() if p.ndim > 1]
    no_decay = [p for p in model.parameters() if p.ndim <= 1]
    return torch.optim.SGD([{'params': decay, 'weight_decay': weight_decay}, {'params': no_decay, 'weight_decay': 0.0}],
                           lr=lr, momentum=momentum)


# --------------------------------------------------------------------------------------------
def cpu_oracle_images_per_sec(model, batch, steps, threads):
    """The reference's CPU path (its model code restated by oracle/) on this box's host cores."""
    from oracle import convnets, train_step, vit
    torch.set_num_threads(threads)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(batch, 3, 224, 224, generator=g)
    y = torch.randint(0, 1000, (batch,), generator=g)
    sd = convnets.init_state('resnet50', 1000, 0) if model == 'resnet50' else vit.init_state(model, 1000, 0)
    buf = {}
    times = []
    for i in range(steps + 1):
        t0 = time.perf_counter()
        if model == 'resnet50':
            _, _, grads = train_step.loss_and_grads(sd, x, y, 'resnet50')
        else:
            _, _, grads = vit.loss_and_grads(sd, x, y, model, global_pool=True)
        train_step.sgd_step(sd, grads, buf, 0.1)  # the CPU cost of the parameter update is negligible either way
        if i > 0:
            times.append(time.perf_counter() - t0)
    dt = sum(times) / len(times)
    return batch / dt, dt


def run_reference(args, rank):
    """--impl reference: the reference's own CPU implementation of the step, all host threads."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    batch = 16
    steps = max(1, min(args.steps, 3))
    ips, dt = cpu_oracle_images_per_sec(args.model, batch, steps, cores)
    line = {
        'impl': 'reference', 'metric': METRICS[args.model], 'value': ips, 'unit': 'images/s', 'n_gpus': args.gpus, 'steps': steps,
        'warmup': 1, 'ms_per_step': dt * 1e3, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': f'{args.model} 224x224 training step, reference model code on CPU',
                   'per_step_batch': batch},
        'cpu_baseline': {'value': ips, 'unit': 'images/s', 'cores': cores, 'kind': 'port',
                         'sample': f'{steps} timed steps of batch {batch} (1 warm-up), fp32, torch {torch.__version__} CPU'},
        'e2e': {'value': ips, 'unit': 'images/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------
class OpTimer:
    """CUDA-event timing of every C-ABI call (installed around _lib.call for an instrumented pass)."""

    def __init__(self):
        self.records = []

    def install(self):
        from simpleaicv_pytorch_training_examples_b200 import _lib
        self._lib = _lib
        self._orig = _lib.call
        timer = self

        def timed(name, *a):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            rc = timer._orig(name, *a)
            e.record()
            timer.records.append((name, a, s, e))
            return rc

        _lib.call = timed
        import simpleaicv_pytorch_training_examples_b200.ops as ops
        ops._lib.call = timed

    def remove(self):
        self._lib.call = self._orig

    def summarize(self, batch):
        torch.cuda.synchronize()
        out = {}
        for name, a, s, e in self.records:
            key, flops, nbytes = describe(name, a)
            d = out.setdefault(key, {'ms': 0.0, 'calls': 0, 'flops': 0.0, 'bytes': 0.0, 'op': name})
            d['ms'] += s.elapsed_time(e)
            d['calls'] += 1
            d['flops'] += flops
            d['bytes'] += nbytes
        return out


def _cs(a):
    cs = a._obj if hasattr(a, '_obj') else a
    return cs


def describe(name, a):
    """(key, algorithmic flops, algorithmic bytes) of one C-ABI call from its arguments."""
    if name in ('saicv_conv_fprop', 'saicv_conv_dgrad', 'saicv_conv_wgrad'):
        cs = next(_cs(v) for v in a if hasattr(v, '_obj'))
        st = cs.stride
        P = (cs.h + 2 * cs.pad - cs.r) // st + 1
        Q = (cs.w + 2 * cs.pad - cs.s) // st + 1
        if name == 'saicv_conv_dgrad':
            P, Q = cs.h, cs.w
        pix = cs.n * P * Q
        flops = 2.0 * pix * cs.k * cs.c * cs.r * cs.s
        x_b, y_b, w_b = cs.n * cs.h * cs.w * cs.c * 2, pix * cs.k * 2, cs.k * cs.c * cs.r * cs.s * 2
        nbytes = x_b + y_b + w_b if name != 'saicv_conv_wgrad' else x_b + y_b + cs.k * cs.c * cs.r * cs.s * 4
        return f'{name[6:]} {cs.r}x{cs.s}/{st} c{cs.c} k{cs.k} {cs.h}x{cs.w}', flops, float(nbytes)
    if name == 'saicv_linear_fwd':      # (..., M, N, K, flags, out_f32, stream)
        M, N, K = a[-6], a[-5], a[-4]
        return f'linear_fwd M{M} N{N} K{K}', 2.0 * M * N * K, 2.0 * (M * K + N * K + M * N)
    if name == 'saicv_linear_dgrad':    # (..., M, N, K, flags, out_f32, stream)
        M, N, K = a[-6], a[-5], a[-4]
        return f'linear_dgrad M{M} N{N} K{K}', 2.0 * M * N * K, 2.0 * (M * K + N * K + M * N)
    if name == 'saicv_linear_wgrad':    # (dy, x, dw, M, N, K, splits, stream)
        M, N, K = a[-5], a[-4], a[-3]
        return f'linear_wgrad M{M} N{N} K{K}', 2.0 * M * N * K, 2.0 * (M * K + M * N) + 4.0 * N * K
    return name[6:], 0.0, 0.0


def run_b200(args, rank, world, local_rank):
    from simpleaicv_pytorch_training_examples_b200 import _lib
    from simpleaicv_pytorch_training_examples_b200.classification import backbones, losses
    from simpleaicv_pytorch_training_examples_b200.distributed import B200DataParallel
    dev = torch.device('cuda', local_rank)
    torch.cuda.set_device(dev)
    B = args.batch
    torch.manual_seed(0)
    g = torch.Generator().manual_seed(1234 + rank)
    x_host = torch.randn(B, 3, 224, 224, generator=g).pin_memory()
    if args.model == 'resnet50':
        model = backbones.resnet50(num_classes=1000).to(dev).train()
        crit = losses.CELoss().to(dev)
        opt = build_optimizer(model)
        y_host = torch.randint(0, 1000, (B,), generator=g).pin_memory()
        workload = 'ResNet-50 224x224 bs256/GPU training step (fwd+CELoss+bwd+grad all-reduce+SGD)'
    else:
        # BASELINE configs[2] / SURVEY.md 8d C3: vit_base_patch16(global_pool, drop_path 0.1), soft labels,
        # OneHotLabelCELoss, AdamW(5e-4, wd .05) with layer-wise lr decay .65
        from simpleaicv_pytorch_training_examples_b200.tools import utils as tutils
        model = backbones.vit_base_patch16(image_size=224, num_classes=1000, drop_path_prob=0.1, global_pool=True).to(dev).train()
        crit = losses.OneHotLabelCELoss().to(dev)

        class _Cfg:
            optimizer = ('AdamW', {'lr': 5e-4, 'global_weight_decay': False, 'weight_decay': 0.05,
                                   'no_weight_decay_layer_name_list': ['position_encoding', 'cls_token'],
                                   'lr_layer_decay': 0.65, 'lr_layer_decay_block': model.blocks, 'block_name': 'blocks'})
        opt, _ = tutils.build_optimizer(_Cfg, model)
        lab = torch.randint(0, 1000, (B,), generator=g)
        one_hot = torch.nn.functional.one_hot(lab, 1000).float() * 0.9 + 0.1 / 1000
        y_host = (0.5 * one_hot + 0.5 * one_hot.roll(1, 0)).pin_memory()
        workload = 'ViT-B/16 224x224 bs256/GPU training step (fwd+OneHotLabelCELoss+bwd+grad all-reduce+AdamW)'
    net = B200DataParallel(model) if world > 1 else model
    x_dev, y_dev = x_host.to(dev), y_host.to(dev)

    def step(x, y):
        loss = crit(net(x), y)
        loss.backward()
        opt.step()
        opt.zero_grad()
        return loss

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local_rank])
        torch.cuda.synchronize()

    def timed(fn, k):
        barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(k):
            fn()
        e.record()
        barrier()
        ms = torch.tensor([s.elapsed_time(e)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    for _ in range(max(3, args.warmup)):
        step(x_dev, y_dev)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = _lib.launch_count()
    ms_total = timed(lambda: step(x_dev, y_dev), args.steps)
    launches = _lib.launch_count() - l0
    ms_step = ms_total / args.steps
    value = B * world / (ms_step / 1e3)

    # end to end through the public API: every step's batch comes from pinned host memory through
    # tools.utils.CudaPrefetcher (the loader wrapper train_classification uses: H2D of batch i+1 on a
    # side stream while step i computes) and the loss is read back to the host every step
    from simpleaicv_pytorch_training_examples_b200.tools.utils import CudaPrefetcher

    def e2e_loop(k):
        host_batches = ({'image': x_host, 'label': y_host} for _ in range(k))
        for batch in CudaPrefetcher(host_batches, dev):
            step(batch['image'], batch['label']).item()

    e2e_loop(2)
    barrier()
    s_, e_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s_.record()
    e2e_loop(args.steps)
    e_.record()
    barrier()
    t_ = torch.tensor([s_.elapsed_time(e_)], device=dev)
    if world > 1:
        dist.all_reduce(t_, op=dist.ReduceOp.MAX)
    e2e_ms = float(t_.item()) / args.steps
    clocks = sampler.stop() if rank == 0 else None
    e2e_value = B * world / (e2e_ms / 1e3)

    roof, cpu_base, table = None, None, None
    peaks = _peaks()
    if rank == 0:
        # instrumented pass (not part of any reported throughput): per-op CUDA-event timing
        # (rank 0 only: the gradient exchange is skipped, otherwise the other ranks would be waited for)
        import contextlib
        t = OpTimer()
        t.install()
        with (net.no_sync() if world > 1 else contextlib.nullcontext()):
            for _ in range(2):
                step(x_dev, y_dev)
        table = t.summarize(B)
        t.remove()
        for d in table.values():
            d['ms'] /= 2
            d['calls'] //= 2
            d['flops'] /= 2
            d['bytes'] /= 2
        gemm = {k: v for k, v in table.items() if v['flops'] > 0}
        # Dominant kernel = gemm_sm100_kernel (every conv / linear fprop, dgrad, wgrad launch of the step).
        # Each launch is bound either by the tensor pipe or by HBM (algorithmic FLOPs and bytes of its
        # shape); per class: achieved = sum(algorithmic work) / sum(CUDA-event durations).
        tf_peak, bw_peak = peaks['tf_sustained'] * 1e12, peaks['hbm_gbs'] * 1e9
        cls = {'tensor': {'work': 0.0, 'ms': 0.0, 'roof_ms': 0.0, 'launches': 0},
               'hbm': {'work': 0.0, 'ms': 0.0, 'roof_ms': 0.0, 'launches': 0}}
        for k, v in gemm.items():
            t_t, t_h = v['flops'] / tf_peak * 1e3, v['bytes'] / bw_peak * 1e3
            c = cls['tensor'] if t_t >= t_h else cls['hbm']
            c['work'] += v['flops'] if t_t >= t_h else v['bytes']
            c['ms'] += v['ms']
            c['roof_ms'] += max(t_t, t_h)
            c['launches'] += v['calls']
            v['roof_frac'] = max(t_t, t_h) / v['ms'] if v['ms'] else 0.0
        dom = 'tensor' if cls['tensor']['ms'] >= cls['hbm']['ms'] else 'hbm'
        d = cls[dom]
        if dom == 'tensor':
            ach, peak, unit = d['work'] / (d['ms'] / 1e3) / 1e12, peaks['tf_sustained'], 'TFLOP/s'
        else:
            ach, peak, unit = d['work'] / (d['ms'] / 1e3) / 1e9, peaks['hbm_gbs'], 'GB/s'
        gemm_ms = sum(v['ms'] for v in gemm.values())
        all_ms = sum(v['ms'] for v in table.values())
        top_key = max(gemm, key=lambda k: gemm[k]['ms'])
        roof = {'bound': dom, 'achieved': ach, 'peak': peak, 'unit': unit, 'frac': ach / peak, 'traffic': None,
                'kernel': 'gemm_sm100_kernel', 'launches_in_class': d['launches'],
                'frac_definition': ('every gemm_sm100_kernel launch of the step is classed tensor- or hbm-bound from its '
                                    'algorithmic FLOPs / bytes vs MEASURED_PEAKS; achieved = sum(work) / sum(CUDA-event time) '
                                    'over the class holding most of the GEMM time; frac = achieved / peak of that class. '
                                    'See other_class, all_launches_frac_of_roofline_time and slowest_launch for the rest.'),
                'definition_changed_from': ('earlier lines of this round reported the single (op, shape) with the largest total '
                                            'time (R50: conv_wgrad 3x3 c64 k64 56x56, frac ~0.145); that launch is still '
                                            'reported under slowest_launch and did not get faster'),
                'peak_source': peaks['source'] + (' (sustained)' if dom == 'tensor' else ''),
                'class_share_of_gemm_time': d['ms'] / gemm_ms,
                'all_launches_frac_of_roofline_time': sum(c['roof_ms'] for c in cls.values()) / gemm_ms,
                'other_class': {k: (c['work'] / (c['ms'] / 1e3) / (1e12 if k == 'tensor' else 1e9) if c['ms'] else None)
                                for k, c in cls.items() if k != dom},
                'slowest_launch': {'launch': top_key, 'ms_per_launch': gemm[top_key]['ms'] / gemm[top_key]['calls'],
                                   'frac_of_its_roofline': gemm[top_key]['roof_frac']},
                'gemm_share_of_step': gemm_ms / all_ms if all_ms else None,
                'step_tensor_tflops': 3 * FWD_FLOPS[args.model] * B / (ms_step / 1e3) / 1e12,
                'step_hbm_gbs_algorithmic': ALGO_BYTES[args.model] * B / (ms_step / 1e3) / 1e9}
        if args.dump_ops:
            rows = sorted(table.items(), key=lambda kv: -kv[1]['ms'])
            with open(args.dump_ops, 'w') as f:
                f.write('op,calls,ms_per_step,GFLOP,algorithmic_MB,TFLOP/s,GB/s,frac_of_roofline\n')
                for k, v in rows:
                    s = v['ms'] / 1e3
                    f.write(f"{k},{v['calls']},{v['ms']:.4f},{v['flops'] / 1e9:.2f},{v['bytes'] / 1e6:.2f},"
                            f"{v['flops'] / s / 1e12 if s else 0:.1f},{v['bytes'] / s / 1e9 if s else 0:.1f},{v.get('roof_frac', 0):.3f}\n")
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            ips, dt = cpu_oracle_images_per_sec(args.model, 8, 2, cores)
            cpu_base = {'value': ips, 'unit': 'images/s', 'cores': cores, 'kind': 'port',
                        'sample': f'2 timed steps of batch 8 (1 warm-up) of the same {args.model} step, fp32 oracle, {dt:.2f} s/step'}

    if rank == 0:
        line = {
            'metric': METRICS[args.model], 'value': value, 'unit': 'images/s', 'n_gpus': world, 'steps': args.steps, 'warmup': max(3, args.warmup),
            'ms_per_step': ms_step, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'bf16',
            'data': 'synthetic',
            'config': {'workload': workload,
                       'per_gpu_batch': B, 'global_batch': B * world, 'parallelism': f'dp{world}',
                       'l2_policy': 'inputs+activations (>10 GB/step) far exceed the 126 MB L2; no explicit flush',
                       'images_per_sec_per_gpu': value / world},
            'clocks': clocks,
            'e2e': {'value': e2e_value, 'unit': 'images/s', 'ms_per_step': e2e_ms,
                    'h2d_bytes_per_step': (x_host.numel() * x_host.element_size() + y_host.numel() * y_host.element_size()) * world, 'd2h_bytes_per_step': 4 * world},
            'gpu_launches': int(launches),
            'roofline': roof,
            'cpu_baseline': cpu_base,
        }
        print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--batch', type=int, default=256, help='per-GPU batch (BASELINE config: 256)')
    ap.add_argument('--model', default='resnet50', choices=['resnet50', 'vit_base_patch16'],
                    help='resnet50 = BASELINE configs[1] (default, the headline); vit_base_patch16 = configs[2]')
    ap.add_argument('--dump-ops', default=None, help='write the per-op timing table (csv) here')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local_rank = int(os.environ.get('LOCAL_RANK', 0))
    if args.impl == 'reference':
        run_reference(args, rank)
        return
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device; the B200 arm has no CPU fallback (use --impl reference for the CPU oracle)')
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        torch.cuda.set_device(local_rank)
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    try:
        run_b200(args, rank, world, local_rank)
    finally:
        if world > 1:
            dist.destroy_process_group()


if __name__ == '__main__':
    main()
