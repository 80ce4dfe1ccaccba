import sys, torch, numpy as np
sys.path.insert(0,'/root/repo')
import models
from fastdepth_b200 import synthetic
from fastdepth_b200.engine import SkipAddEngine
from oracle import fastdepth_oracle as orc
def rel(got,want):
    got=got.double(); want=want.double(); d=torch.maximum(want.abs(), want.abs().mean()); return ((got-want).abs()/d).max().item()
def check(widths,n,h,w,tma,tag,upto=8):
    sd=synthetic.synthetic_state_dict(widths)
    m=models.MobileNetSkipAdd((h,w),pretrained=False,widths=widths); m.load_state_dict(sd); m=m.eval().cuda().half()
    eng=SkipAddEngine(m); eng.set_option('graph',0); eng.set_option('tma_epilogue',tma); eng.set_option('inplace_skip',0)
    m.__dict__['_fd_engine']=eng
    x=synthetic.synthetic_input(n,h,w,seed=4)
    with torch.no_grad(): y=m(x.cuda().half())
    torch.cuda.synchronize()
    sdq={k:(v.half().float() if v.is_floating_point() else v) for k,v in sd.items()}
    st={}; want=orc.skipadd_forward(sdq,x.half().float(),stages=st)
    plan=next(iter(eng.plans.values()))
    ks={s['stage']:s['kernel'] for s in plan.steps()}
    out=[]
    for i,name in enumerate(plan.names[:upto]):
        got=plan.stage_tensor(i).float().cpu().permute(0,3,1,2); ref=st[name]
        e=rel(got,ref); out.append('%s %.3g'%(name,e))
        if e>0.1:
            d=(got-ref).abs(); bad=(d>0.2*ref.abs().mean())
            print('   BAD',name,ks[i],'bad channels:',sorted(set(bad.nonzero()[:,1].tolist()))[:40],'bad frac %.3f'%bad.float().mean())
            break
    print(tag,' '.join(out))
E=list(synthetic.PRUNED_ENCODER); D=synthetic.PRUNED_DECODER
for c4 in (144,128,160,192,200):
    e=list(E); e[4]=c4
    check((tuple(e),D),3,96,64,1,'c4=%d'%c4)
