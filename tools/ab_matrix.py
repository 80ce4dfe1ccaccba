"""A/B a list of (environment knobs, plan options) on the bench workload: whole-forward graph time and per-stage times
(L2 flushed).  usage: python tools/ab_matrix.py  [stock|pruned] ['ENV=1,opt=2' ...]   (no configs: the built-in matrix)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import models  # noqa: E402
from fastdepth_b200 import synthetic  # noqa: E402
from fastdepth_b200.engine import SkipAddEngine  # noqa: E402

args = sys.argv[1:]
widths = synthetic.PRUNED_WIDTHS if (args and args[0] == 'pruned') else synthetic.STOCK_WIDTHS
if args and args[0] in ('stock', 'pruned'):
    args = args[1:]
configs = args or ['', 'FD_TC_NO_HALFK=1', 'wait_sleep_ns=200', 'wait_sleep_ns=500', 'wait_sleep_ns=1000',
                   'wait_sleep_ns=500,FD_TC_MMA_SLEEP=100', 'wait_sleep_ns=500,FD_TC_MMA_SLEEP=100,FD_TC_DW_SLEEP=100',
                   'wait_sleep_ns=2000,FD_TC_MMA_SLEEP=200']
sd = synthetic.synthetic_state_dict(widths)
m = models.MobileNetSkipAdd((224, 224), pretrained=False, widths=widths)
m.load_state_dict(sd)
m = m.eval().cuda().half()
x = synthetic.synthetic_input(64, 224, 224).cuda().half()
y = torch.empty((64, 1, 224, 224), dtype=torch.half, device='cuda')
sp = torch.cuda.current_stream().cuda_stream
ENVK = [k for c in configs for k in (kv.split('=')[0] for kv in c.split(',') if kv) if k.isupper()]
ref = None
for c in configs:
    for k in ENVK:
        os.environ.pop(k, None)
    opts = {}
    for kv in (c.split(',') if c else []):
        k, v = kv.split('=')
        if k.isupper():
            os.environ[k] = v
        else:
            opts[k] = int(v)
    eng = SkipAddEngine(m)
    for k, v in opts.items():
        eng.set_option(k, v)
    plan = eng.plan_for(x)
    for _ in range(5):
        plan.forward(x, y, sp)
    torch.cuda.synchronize()
    if ref is None:
        ref = y.clone()
    same = bool(torch.equal(ref, y))
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50):
            plan.forward(x, y, sp)
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / 50)
    t = plan.time_steps(x, y, sp, warmup=2, iters=8, flush_l2=True)
    print('[%s] forward %.1f us (%.0f img/s) bits_equal_first=%s sum_stages %.0f' % (c or 'default', best * 1e3, 64 / best * 1e3, same, sum(s['ms'] for s in t) * 1e3))
    print('    ' + ' '.join('%s:%.1f' % (s['stage_name'].replace('decode_conv', 'd').replace('conv', 'c'), s['ms'] * 1e3) for s in t), flush=True)
    eng.refresh()
