"""Bring-up helper for the block kernel's cluster mode: run one small forward stage by stage (graph off, chain off) with a
forced cluster size and compare every stage output with the cluster-less plan.  usage: python tools/cluster_dbg.py [cs] [n h w]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import models
from fastdepth_b200 import synthetic
from fastdepth_b200.engine import SkipAddEngine
cs = sys.argv[1] if len(sys.argv) > 1 else '4'
n, h, w = (int(a) for a in sys.argv[2:5]) if len(sys.argv) > 4 else (2, 64, 96)
widths = synthetic.PRUNED_WIDTHS if os.environ.get('WIDTHS') == 'pruned' else synthetic.STOCK_WIDTHS
sd = synthetic.synthetic_state_dict(widths)
m = models.MobileNetSkipAdd((h, w), pretrained=False, widths=widths); m.load_state_dict(sd); m = m.eval().cuda().half()
x = synthetic.synthetic_input(n, h, w, seed=3).cuda().half()
outs = {}
for tag, env in (('ref', '1'), ('cl', cs)):
    if env.startswith('w'):                       # 'w2' / 'w4': weight-multicast clusters instead of tile-sharing ones
        os.environ['FD_TC_CLUSTER'] = '1'; os.environ['FD_TC_WMC'] = env[1:]
    else:
        os.environ['FD_TC_CLUSTER'] = env; os.environ['FD_TC_WMC'] = '1'
    eng = SkipAddEngine(m)
    for k, v in (('graph', 0), ('chain', 0), ('inplace_skip', 0), ('fold_head', 0), ('pdl', int(os.environ.get('PDL', '1')))):
        eng.set_option(k, v)
    plan = eng.plan_for(x)
    y = torch.empty((n, 1, h, w), dtype=torch.half, device='cuda')
    sp = torch.cuda.current_stream().cuda_stream
    for s in plan.steps():
        print(tag, s['stage'], s['kernel'], flush=True)
    plan.forward(x, y, sp)
    torch.cuda.synchronize()
    outs[tag] = [plan.stage_tensor(i).clone() for i in range(len(plan.names) - 1)] + [y.clone()]
    print(tag, 'forward done', flush=True)
bad = 0
for i, (a, b) in enumerate(zip(outs['ref'], outs['cl'])):
    d = (a.float() - b.float()).abs().max().item()
    print('stage %2d max abs diff %g%s' % (i, d, '' if d == 0 else '   <-- DIFFERS'))
    bad += d != 0
print('OK' if not bad else 'MISMATCH in %d stages' % bad)
