import sys, torch, numpy as np
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
import models
from fastdepth_b200 import synthetic
from fastdepth_b200.engine import SkipAddEngine
from oracle import fastdepth_oracle as orc
def rel(got,want):
    got=got.double(); want=want.double(); d=torch.maximum(want.abs(), want.abs().mean()); return ((got-want).abs()/d).max().item()
def check(widths,n,h,w,tma,inpl,tag):
    sd=synthetic.synthetic_state_dict(widths)
    m=models.MobileNetSkipAdd((h,w),pretrained=False,widths=widths); m.load_state_dict(sd); m=m.eval().cuda().half()
    eng=SkipAddEngine(m); eng.set_option('graph',0); eng.set_option('tma_epilogue',tma); eng.set_option('inplace_skip',inpl); eng.set_option('fold_head',1)
    m.__dict__['_fd_engine']=eng
    x=synthetic.synthetic_input(n,h,w,seed=4)
    with torch.no_grad(): y=m(x.cuda().half())
    torch.cuda.synchronize()
    sdq={k:(v.half().float() if v.is_floating_point() else v) for k,v in sd.items()}
    st={}; want=orc.skipadd_forward(sdq,x.half().float(),stages=st)
    plan=next(iter(eng.plans.values()))
    out=[]
    for i,name in enumerate(plan.names[:-1]):
        if name=='decode_conv5': continue
        got=plan.stage_tensor(i).float().cpu().permute(0,3,1,2)
        ref=st[name]
        if inpl and name in ('conv1','conv3','conv5'): ref=st[{'conv1':'decode_conv4','conv3':'decode_conv3','conv5':'decode_conv2'}[name]]
        e=rel(got,ref)
        # per-image error to see which images break
        pe=[rel(got[j:j+1],ref[j:j+1]) for j in range(got.shape[0])]
        out.append('%s %.3g%s'%(name,e,'' if e<0.06 else ' imgs:'+','.join('%.2g'%v for v in pe)))
    print(tag,'tma',tma,'inpl',inpl,'| final %.3g |'%rel(y.float().cpu(),want),' '.join(out))
check(synthetic.PRUNED_WIDTHS,3,96,64,1,0,'pruned3')
check(synthetic.STOCK_WIDTHS,8,224,224,0,0,'stock8')
check(synthetic.STOCK_WIDTHS,8,224,224,1,0,'stock8')
check(synthetic.STOCK_WIDTHS,8,224,224,1,1,'stock8')
