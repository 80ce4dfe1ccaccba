"""One forward of the metric batch (64x224x224 fp16, chain off unless CHAIN=1) under the planner knobs of the environment; prints
the kernels and a checksum.  usage: [ENV=...] python tools/cfg_run.py [stock|pruned] [opt=val ...]"""
import hashlib, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import models
from fastdepth_b200 import synthetic
from fastdepth_b200.engine import SkipAddEngine
args = sys.argv[1:]
widths = synthetic.PRUNED_WIDTHS if (args and args[0] == 'pruned') else synthetic.STOCK_WIDTHS
opts = dict(kv.split('=') for kv in args if '=' in kv)
sd = synthetic.synthetic_state_dict(widths)
m = models.MobileNetSkipAdd((224, 224), pretrained=False, widths=widths); m.load_state_dict(sd); m = m.eval().cuda().half()
x = synthetic.synthetic_input(64, 224, 224, seed=11).cuda().half()
eng = SkipAddEngine(m)
eng.set_option('chain', int(os.environ.get('CHAIN', '0')))
for k, v in opts.items():
    eng.set_option(k, int(v))
m.__dict__['_fd_engine'] = eng
with torch.no_grad():
    y = m(x).clone()
    y2 = m(x).clone()
torch.cuda.synchronize()
ks = ' '.join(s['kernel'].split('[')[0].replace('block_tc<', '<') for s in next(iter(eng.plans.values())).steps())
print(hashlib.md5(y.cpu().numpy().tobytes()).hexdigest(), bool(torch.equal(y, y2)), ks, flush=True)
