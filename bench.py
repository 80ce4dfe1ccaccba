#!/usr/bin/env python
"""bench.py -- depth-maps/sec of the FastDepth forward path (MobileNetSkipAdd.forward).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one pass of the hot path over one batch of synthetic input: batch 64 per GPU,
224x224, fp16, MobileNet-NNConv5(dw)+skipadd (BASELINE.json metric config; weak scaling: every
rank keeps 64 images).  Prints ONE JSON line on rank 0.

  value      whole-job images/s with inputs resident in HBM (CUDA events, max over ranks)
  e2e        same metric through the C-ABI host-buffer call fd_forward_host: pinned-host -> device
             copy of the batch and device -> host copy of the depth maps inside the timed region
  roofline   the dominant kernel's achieved HBM GB/s = algorithmic bytes / launch duration
             (each kernel timed alone with CUDA events, L2 flushed between launches)
  cpu_baseline  the oracle port of the reference forward (torch CPU fp32) on this box's host cores,
             bounded sample
  --impl reference : times ONLY that CPU implementation (the reference is pure Python on PyTorch;
             /root/reference does not exist on the GPU box, so the oracle port stands in).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = 'depth-maps/sec @224x224 b64'
UNIT = 'images/s'


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=30)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--batch', type=int, default=64, help='images per GPU')
    ap.add_argument('--hw', type=int, nargs=2, default=[224, 224])
    ap.add_argument('--dtype', default='fp16', choices=['fp16', 'bf16', 'fp32'])
    ap.add_argument('--widths', default='stock', choices=['stock', 'pruned'])
    ap.add_argument('--path', type=int, default=1)
    ap.add_argument('--fold-head', type=int, default=1)
    ap.add_argument('--graph', type=int, default=1)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--stage-iters', type=int, default=10)
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs, burst copy)'
    return 6650.0, 'fallback (B200_PROFILING.md)'


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons during the timed region (recipe's clocks line)."""
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                      '--format=csv,noheader,nounits'], capture_output=True, text=True, timeout=5).stdout
                f = [s.strip() for s in out.strip().split(',')]
                if len(f) >= 7:
                    self.samples.append(f)
            except Exception:
                pass
            time.sleep(0.05)

    def summary(self):
        self.stop_flag = True
        self.join(timeout=6)
        if not self.samples:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        sm = sorted(float(s[0]) for s in self.samples)
        reasons = set()
        for s in self.samples:
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), s[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        return {'sm_mhz': sm[len(sm) // 2], 'sm_max_mhz': float(self.samples[0][1]), 'reasons': sorted(reasons),
                'samples': len(self.samples), 'power_w_max': max(float(s[2]) for s in self.samples)}


def build_sd(widths_name):
    from fastdepth_b200 import synthetic
    widths = synthetic.STOCK_WIDTHS if widths_name == 'stock' else synthetic.PRUNED_WIDTHS
    return widths, synthetic.synthetic_state_dict(widths, seed=1)


def usable_cpus():
    """CPUs this process may really use: the affinity mask, capped by the cgroup CPU quota (a container that shows 128
    cores but is throttled to a few turns a 128-thread OpenMP run into seconds per image)."""
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        n = os.cpu_count() or 1
    for path in ('/sys/fs/cgroup/cpu.max', '/sys/fs/cgroup/cpu/cpu.cfs_quota_us'):
        try:
            txt = open(path).read().split()
            if path.endswith('cpu.max'):
                if txt[0] != 'max':
                    n = min(n, max(1, int(float(txt[0]) / float(txt[1]) + 0.5)))
            else:
                q = int(txt[0])
                if q > 0:
                    per = int(open('/sys/fs/cgroup/cpu/cpu.cfs_period_us').read())
                    n = min(n, max(1, int(q / per + 0.5)))
            break
        except Exception:
            continue
    return max(1, n)


def pick_cpu_threads(sd, h, w):
    """The reference arm is meant to use all the host threads it can use WELL: sweep the thread count upwards on a
    one-image forward and keep the fastest (stops as soon as more threads make it slower).  Returns (threads, note)."""
    from fastdepth_b200 import synthetic
    from oracle import fastdepth_oracle as orc          # the CPU baseline leg may execute the oracle
    torch.set_grad_enabled(False)
    top = usable_cpus()
    cands = sorted({c for c in (4, 8, 16, 32, 64, 128, 256, top // 2, top) if 1 <= c <= top})
    x1 = synthetic.synthetic_input(1, h, w, seed=0)
    best_t, best_c, log = None, cands[0], []
    for c in cands:
        torch.set_num_threads(c)
        orc.skipadd_forward(sd, x1)
        t0 = time.perf_counter(); orc.skipadd_forward(sd, x1); t = time.perf_counter() - t0
        log.append('%d:%.0fms' % (c, t * 1e3))
        if best_t is None or t < best_t:
            best_t, best_c = t, c
        elif t > 1.5 * best_t:
            break
    torch.set_num_threads(best_c)
    return best_c, 'threads swept on a 1-image forward (%s of %d usable CPUs)' % (' '.join(log), top)


def cpu_forward_rate(sd, h, w, budget_s, min_steps, warmup, batch=None):
    """Time the oracle port of the reference forward on the host cores; returns (img/s, batch, steps, s/step)."""
    from fastdepth_b200 import synthetic
    from oracle import fastdepth_oracle as orc          # the CPU baseline leg may execute the oracle
    torch.set_grad_enabled(False)
    if batch is None:
        x1 = synthetic.synthetic_input(1, h, w, seed=0)
        orc.skipadd_forward(sd, x1)
        t0 = time.perf_counter(); orc.skipadd_forward(sd, x1); t1 = time.perf_counter() - t0
        batch = max(1, min(8, int(1.0 / max(t1, 1e-3))))   # ~1 s of work per step, main.py feeds bs 1 (l.41)
    x = synthetic.synthetic_input(batch, h, w, seed=0)
    for _ in range(warmup):
        orc.skipadd_forward(sd, x)
    times = []
    t_start = time.perf_counter()
    while len(times) < min_steps or (budget_s and time.perf_counter() - t_start < budget_s and len(times) < 10 * min_steps):
        t0 = time.perf_counter()
        orc.skipadd_forward(sd, x)
        times.append(time.perf_counter() - t0)
        if budget_s and time.perf_counter() - t_start > budget_s and len(times) >= min_steps:
            break
    per = sum(times) / len(times)
    return batch / per, batch, len(times), per


def run_reference(args, rank):
    if rank != 0:
        return
    widths, sd = build_sd(args.widths)
    h, w = args.hw
    # torchrun exports OMP_NUM_THREADS=1 for every worker; the reference arm is meant to use all host threads it can
    cores, thread_note = pick_cpu_threads(sd, h, w)
    rate, batch, steps, per = cpu_forward_rate(sd, h, w, budget_s=0, min_steps=max(1, args.steps), warmup=max(1, args.warmup))
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': rate, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': steps,
        'warmup': max(1, args.warmup), 'ms_per_step': per * 1e3, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': 'MobileNet-NNConv5(dw)+skipadd %s widths, %dx%d, reference forward on host CPU' %
                               (args.widths, h, w), 'batch_per_step': batch, 'global_batch': batch},
        'cpu_baseline': {'value': rate, 'unit': UNIT, 'cores': cores, 'kind': 'port',
                         'sample': '%d steps of batch %d (fp32, torch CPU, NCHW) -- the reference is pure Python on '
                                   'PyTorch and /root/reference is absent on the GPU box, so the oracle port runs; %s' % (steps, batch, thread_note)},
        'e2e': {'value': rate, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)


def main():
    args = parse()
    rank = int(os.environ.get('RANK', 0))
    local_rank = int(os.environ.get('LOCAL_RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    if args.impl == 'reference':
        run_reference(args, rank)
        return

    import torch.distributed as dist
    import models
    from fastdepth_b200 import synthetic
    from fastdepth_b200.engine import SkipAddEngine

    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a GPU (no CPU fallback); use --impl reference for the CPU arm')
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    dtype = {'fp16': torch.float16, 'bf16': torch.bfloat16, 'fp32': torch.float32}[args.dtype]
    h, w = args.hw
    n = args.batch
    widths, sd = build_sd(args.widths)
    model = models.MobileNetSkipAdd((h, w), pretrained=False, widths=widths)
    model.load_state_dict(sd)
    model = model.eval().to(dev).to(dtype)
    eng = SkipAddEngine(model)
    for k, v in (('path', args.path), ('fold_head', args.fold_head), ('graph', args.graph)):
        eng.set_option(k, v)
    model.__dict__['_fd_engine'] = eng

    # 4 rotating input batches (different images per rank); a step moves >1 GB through HBM, far more
    # than the 126 MB L2, so nothing of the previous step's input survives in cache.
    n_rot = 4
    xs = [synthetic.synthetic_input(n, h, w, seed=100 * rank + i).to(dev).to(dtype) for i in range(n_rot)]
    y = torch.empty((n, 1, h, w), dtype=dtype, device=dev)
    plan = eng.plan_for(xs[0])
    stream = torch.cuda.current_stream(dev)
    sp = stream.cuda_stream

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn, steps, warmup):
        for i in range(warmup):
            fn(i)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for i in range(steps):
            fn(i)
        e1.record(stream)
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    # ---- value: device-resident inputs ---------------------------------------------------------
    ms_total = timed(lambda i: plan.forward(xs[i % n_rot], y, sp), args.steps, max(3, args.warmup))
    # ---- e2e: pinned host buffers through the C-ABI pipeline (fd_pipeline_submit / fd_pipeline_wait): every step
    # uploads its batch from pinned host memory and downloads its depth maps; up to 3 batches are in flight so
    # the PCIe copies overlap the forward of the neighbouring steps.  Timed on the host clock between device-wide
    # synchronisations (the work spans three streams), max over ranks.
    xh = [x.cpu().pin_memory() for x in xs[:3]]
    yh = [torch.empty((n, 1, h, w), dtype=dtype).pin_memory() for _ in range(3)]
    e2e_steps = max(6, args.steps)

    def run_pipeline(k):
        tickets = []
        for i in range(k):
            tickets.append(plan.pipeline_submit(xh[i % 3], yh[i % 3]))
            if i >= 2:
                plan.pipeline_wait(tickets[i - 2])
        for t in tickets[-2:]:
            plan.pipeline_wait(t)

    run_pipeline(4)
    barrier()
    t0 = time.perf_counter()
    run_pipeline(e2e_steps)
    torch.cuda.synchronize(dev)
    el = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(el, op=dist.ReduceOp.MAX)
    ms_e2e = el.item() * 1e3
    # the plain synchronous call, for reference
    ms_sync = timed(lambda i: plan.forward_host(xh[i % 3], yh[i % 3], sp), 5, 2) / 5
    clocks = sampler.summary() if sampler else None

    value = world * n * args.steps / (ms_total * 1e-3)
    e2e_value = world * n * e2e_steps / (ms_e2e * 1e-3)

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- per-kernel roofline (rank 0, each kernel alone, L2 flushed) ---------------------------
    hbm_peak, peak_src = peaks()
    steps = plan.time_steps(xs[0], y, sp, warmup=2, iters=args.stage_iters, flush_l2=True)
    for s in steps:
        s['gbs'] = s['alg_bytes'] / (s['ms'] * 1e-3) / 1e9 if s['ms'] > 0 else 0.0
        s['frac'] = s['gbs'] / hbm_peak
        s['tflops'] = 2 * s['macs'] / (s['ms'] * 1e-3) / 1e12 if s['ms'] > 0 else 0.0
    top = max(steps, key=lambda s: s['ms'])
    # DRAM traffic of that kernel from the committed `ncu --set full` capture of the same configuration (if any)
    traffic = None
    tpath = os.path.join(ROOT, 'profiles', 'r01_final_traffic.json')
    if os.path.exists(tpath) and args.widths == 'stock' and args.dtype == 'fp16' and n == 64 and (h, w) == (224, 224) and args.path == 1:
        tj = json.load(open(tpath))['stages'].get(top['stage_name'])
        if tj:
            traffic = (tj['dram_read_mb'] + tj['dram_write_mb']) * 1e6
    sum_ms = sum(s['ms'] for s in steps)
    alg_total = sum(s['alg_bytes'] for s in steps)

    # ---- parity on this very configuration (2 images vs the oracle) ------------------------------
    from oracle import fastdepth_oracle as orc           # checker only
    with torch.no_grad():
        got = model(xs[0][:2].clone()).float().cpu()
    sdq = {k: (v.to(dtype).float() if v.is_floating_point() else v) for k, v in sd.items()}
    want = orc.skipadd_forward(sdq, xs[0][:2].float().cpu())
    denom = torch.maximum(want.abs(), want.abs().mean())
    max_rel = ((got - want).abs() / denom).max().item()
    tgt = synthetic.synthetic_target(want, seed=1)
    m_ours, _ = orc.average_per_image(got.numpy(), tgt.numpy())
    m_ref, _ = orc.average_per_image(want.numpy(), tgt.numpy())

    cpu = None
    if not args.no_cpu_baseline and world == 1:
        ccores, thread_note = pick_cpu_threads(sd, h, w)
        rate, cb, csteps, per = cpu_forward_rate(sd, h, w, budget_s=15.0, min_steps=3, warmup=1)
        cpu = {'value': rate, 'unit': UNIT, 'cores': ccores, 'kind': 'port',
               'sample': '%d forwards of batch %d at %dx%d, fp32 torch CPU (oracle port of reference models.py:706-732), '
                         '%.2f s each; %s' % (csteps, cb, h, w, per, thread_note)}

    line = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
        'warmup': max(3, args.warmup), 'ms_per_step': ms_total / args.steps, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': {'fp16': 'f16', 'bf16': 'bf16', 'fp32': 'f32'}[args.dtype],
        'data': 'synthetic',
        'config': {'workload': 'MobileNet-NNConv5(dw)+skipadd (%s widths) forward, batch %d/GPU, %dx%d, %s'
                               % (args.widths, n, h, w, args.dtype),
                   'global_batch': n * world, 'parallelism': 'image-sharded x%d (no data-path collective)' % world,
                   'path': args.path, 'fold_head': args.fold_head, 'graph': args.graph,
                   'l2': '4 rotating input batches; one step streams %.2f GB through HBM (>> 126 MB L2)' % (alg_total / 1e9),
                   'weights': 'random-init (synthetic recipe seed 1)'},
        'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': xh[0].numel() * xh[0].element_size(),
                'd2h_bytes_per_step': yh[0].numel() * yh[0].element_size(), 'ms_per_step': ms_e2e / e2e_steps,
                'api': 'fd_pipeline_submit/fd_pipeline_wait (C-ABI, pinned host buffers, 3 batches in flight)',
                'sync_call_ms_per_step': ms_sync, 'sync_call_api': 'fd_forward_host'},
        'gpu_launches': plan.launches_per_forward() * args.steps,
        'launches_per_step': plan.launches_per_forward(),
        'clocks': clocks,
        'roofline': {'bound': 'hbm', 'achieved': top['gbs'], 'peak': hbm_peak, 'unit': 'GB/s', 'frac': top['frac'],
                     'traffic': traffic, 'kernel': top['kernel'], 'stage': top['stage_name'], 'peak_source': peak_src,
                     'kernel_ms': top['ms'], 'share_of_step': top['ms'] / sum_ms,
                     'whole_step': {'alg_bytes': alg_total, 'gbs_at_value': alg_total / (ms_total / args.steps * 1e-3) / 1e9,
                                    'frac_at_value': alg_total / (ms_total / args.steps * 1e-3) / 1e9 / hbm_peak}},
        'stages': [{'stage': s['stage_name'], 'kernel': s['kernel'], 'ms': round(s['ms'], 5),
                    'alg_mb': round(s['alg_bytes'] / 1e6, 3), 'gbs': round(s['gbs'], 1), 'frac': round(s['frac'], 4),
                    'tflops': round(s['tflops'], 2)} for s in steps],
        'parity': {'max_rel_err_vs_oracle': max_rel, 'delta1': m_ours['delta1'], 'delta1_oracle': m_ref['delta1'],
                   'rmse_mm': m_ours['rmse'], 'rmse_mm_oracle': m_ref['rmse']},
        'cpu_baseline': cpu,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
