"""Per-stage times with the L2 flushed before every launch vs left warm (same launch repeated), bench workload."""
import sys, os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import models
from fastdepth_b200 import synthetic
from fastdepth_b200.engine import SkipAddEngine
sd = synthetic.synthetic_state_dict()
m = models.MobileNetSkipAdd((224, 224), pretrained=False); m.load_state_dict(sd); m = m.eval().cuda().half()
x = synthetic.synthetic_input(64, 224, 224).cuda().half()
y = torch.empty((64, 1, 224, 224), dtype=torch.half, device='cuda')
sp = torch.cuda.current_stream().cuda_stream
eng = SkipAddEngine(m); plan = eng.plan_for(x)
for _ in range(3): plan.forward(x, y, sp)
torch.cuda.synchronize()
for flush in (True, False):
    t = plan.time_steps(x, y, sp, warmup=2, iters=10, flush_l2=flush)
    print('flush' if flush else 'warm ', ' '.join(f"{s['stage_name'].replace('decode_conv','d').replace('conv','c')}:{s['ms']*1e3:.1f}" for s in t),
          ' sum %.1f' % sum(s['ms'] * 1e3 for s in t))
