import sys, torch
sys.path.insert(0,'/root/repo')
import models
from fastdepth_b200 import synthetic
from fastdepth_b200.engine import SkipAddEngine
sd=synthetic.synthetic_state_dict()
m=models.MobileNetSkipAdd((64,96),pretrained=False); m.load_state_dict(sd); m=m.eval().cuda().half()
eng=SkipAddEngine(m); eng.set_option('graph',0); m.__dict__['_fd_engine']=eng
x=synthetic.synthetic_input(2,64,96).cuda().half()
plan=eng.plan_for(x)
y=torch.empty((2,1,64,96),dtype=torch.half,device='cuda')
try:
    st=plan.time_steps(x,y,torch.cuda.current_stream().cuda_stream,warmup=0,iters=1,flush_l2=False)
    for s in st: print(s['stage_name'],s['kernel'],s['ms'])
except Exception as e:
    print('ERR',e)
