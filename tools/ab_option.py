"""A/B one plan option on the bench workload: per-stage times (L2 flushed) and whole-forward graph time.
usage: python tools/ab_option.py dw_helpers 0 1 [--trace 7 18]"""
import sys, os, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import models
from fastdepth_b200 import synthetic
from fastdepth_b200.engine import SkipAddEngine

args = sys.argv[1:]
trace = []
if '--trace' in args:
    i = args.index('--trace'); trace = [int(a) for a in args[i + 1:]]; args = args[:i]
name, values = args[0], [int(v) for v in args[1:]]
sd = synthetic.synthetic_state_dict()
m = models.MobileNetSkipAdd((224, 224), pretrained=False); m.load_state_dict(sd); m = m.eval().cuda().half()
x = synthetic.synthetic_input(64, 224, 224).cuda().half()
y = torch.empty((64, 1, 224, 224), dtype=torch.half, device='cuda')
sp = torch.cuda.current_stream().cuda_stream
for v in values:
    eng = SkipAddEngine(m); eng.set_option(name, v)
    plan = eng.plan_for(x)
    for _ in range(5): plan.forward(x, y, sp)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50): plan.forward(x, y, sp)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 50
    print(f'{name}={v}: forward {ms*1e3:.1f} us  ({64/ms*1e3:.0f} img/s)')
    t = plan.time_steps(x, y, sp, warmup=2, iters=5, flush_l2=True)
    print('   ', ' '.join(f"{s['stage_name'].replace('decode_conv','d').replace('conv','c')}:{s['ms']*1e3:.0f}" for s in t))
    if trace:
        eng.set_option('graph', 0); plan = eng.plan_for(x); plan.forward(x, y, sp); torch.cuda.synchronize()
        for st in trace:
            tr = plan.trace_stage(st, y, sp)
            t0 = min(a.min() for a in tr.values() if len(a))
            print('  == stage', st, plan.names[st])
            for k, a in tr.items():
                a = a - t0
                print('   %-16s n=%3d %s ... %s' % (k, len(a), ' '.join('%6d' % q for q in a[:12]), ' '.join('%6d' % q for q in a[-3:])))
