import sys, torch, numpy as np
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
import models
from fastdepth_b200 import synthetic
from fastdepth_b200.engine import SkipAddEngine
sd=synthetic.synthetic_state_dict()
m=models.MobileNetSkipAdd((224,224),pretrained=False); m.load_state_dict(sd); m=m.eval().cuda().half()
eng=SkipAddEngine(m); eng.set_option('graph',0); eng.set_option('chain',int(__import__('os').environ.get('CHAIN','1'))); m.__dict__['_fd_engine']=eng
x=synthetic.synthetic_input(64,224,224).cuda().half()
plan=eng.plan_for(x)
y=torch.empty((64,1,224,224),dtype=torch.half,device='cuda')
sp=torch.cuda.current_stream().cuda_stream
plan.forward(x,y,sp); torch.cuda.synchronize()
for st in [int(a) for a in sys.argv[1:]]:
    tr=plan.trace_stage(st,y,sp)
    t0=min(v.min() for v in tr.values() if len(v))
    print('== stage',st,plan.names[st],[s['kernel'] for s in plan.steps() if s['stage']==st])
    for k,v in tr.items():
        v=v-t0
        print('%-13s n=%3d first %s'%(k,len(v),' '.join('%6d'%a for a in v[:14])), ' ... last', ' '.join('%6d'%a for a in v[-3:]))
    if len(tr['dw_start'])>4:
        d=np.diff(tr['dw_start']); print('dw_start period: median %d  mean %d'%(np.median(d),d.mean()))
        print('dw math (start->done) median', np.median(tr['dw_math_done']-tr['dw_start']), ' publish wait+write median', np.median(tr['a_published']-tr['dw_math_done']))
        e=tr['epi_done']-tr['epi_start']; print('epilogue duration median',np.median(e), 'epi period median', np.median(np.diff(tr['epi_start'])) if len(e)>1 else -1)
        print('mma ready->issued median', np.median(tr['mma_issued']-tr['mma_ready']))
        if len(tr['epi_tmem_loaded'])>2:
            n=min(len(tr['epi_start']),len(tr['epi_tmem_loaded']),len(tr['epi_staged']),len(tr['epi_barrier']),len(tr['epi_store_issued']))
            a=[tr[k][:n] for k in ('epi_start','epi_tmem_loaded','epi_staged','epi_barrier','epi_store_issued','epi_done')]
            print('epilogue phases (median cycles): ld %d | math+sts+fence %d | wait_read+barrier %d | store issue %d | rest %d'%tuple(np.median(a[i+1]-a[i]) for i in range(5)))
            print('acc_full wait: epi_start - prev epi_done median', np.median(tr['epi_start'][1:n]-tr['epi_done'][:n-1]))
