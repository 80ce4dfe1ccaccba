"""Throughput of R independent forwards in flight (R plans with their own intermediates, R streams, round-robin) against the
single-stream graph replay: do another batch's kernels fill the ~6 us bubbles at every kernel boundary?
usage: python tools/overlap_test.py [pdl]"""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import models
from fastdepth_b200 import synthetic
from fastdepth_b200.engine import SkipAddEngine
pdl = int(sys.argv[1]) if len(sys.argv) > 1 else 1
sd = synthetic.synthetic_state_dict(synthetic.STOCK_WIDTHS)
m = models.MobileNetSkipAdd((224, 224), pretrained=False); m.load_state_dict(sd); m = m.eval().cuda().half()
xs = [synthetic.synthetic_input(64, 224, 224, seed=i).cuda().half() for i in range(4)]
for R in ((1, 2, 3, 4, 6) if len(sys.argv) > 2 else (1, 2, 3)):
    engs, plans, streams, ys = [], [], [], []
    for r in range(R):
        e = SkipAddEngine(m); e.set_option('pdl', pdl)
        engs.append(e); plans.append(e.plan_for(xs[0])); streams.append(torch.cuda.Stream())
        ys.append(torch.empty((64, 1, 224, 224), dtype=torch.half, device='cuda'))
    torch.cuda.synchronize()
    def run(k):
        for i in range(k):
            r = i % R
            plans[r].forward(xs[i % 4], ys[r], streams[r].cuda_stream)
    run(6 * R); torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        run(120); torch.cuda.synchronize()
        best = min(best, (time.perf_counter() - t0) / 120)
    print('pdl=%d  %d forward(s) in flight: %.1f us per batch  (%.0f img/s)' % (pdl, R, best * 1e6, 64 / best), flush=True)
    ref = ys[0].clone() if R == 1 else ref
    if R > 1:
        run(R); torch.cuda.synchronize()
    del engs, plans
