import sys, time, torch
sys.path.insert(0,'alphazero-general_b200')
from azb200 import nnet as aznet
dev=torch.device('cuda')
def bench(fn, n=20):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/n
for netname, na in [('default', aznet.DEFAULT_NET_ARGS), ('c4train', aznet.CONNECT4_TRAIN_NET_ARGS)]:
  for B in [4096, 8192]:
    for prec in ['fp32','tf32','bf16']:
      for cl in [False, True]:
        for bm in [False, True]:
            torch.backends.cudnn.benchmark=bm
            torch.manual_seed(0)
            m=aznet.ResNet((4,6,7),7,3,**na).to(dev).eval()
            obs=torch.rand(B,4,6,7,device=dev); pol=torch.empty(B,7,device=dev); val=torch.empty(B,3,device=dev)
            try:
                ev=aznet.LeafEvaluator(m,obs,pol,val,precision=prec,use_graph=True,channels_last=cl)
                t=bench(lambda: ev())
                print(f'{netname} B={B} {prec} cl={cl} bench={bm}: {t*1000:.0f} us  -> {B/t/1e3:.2f} M evals/s', flush=True)
            except Exception as ex:
                print(netname,B,prec,cl,bm,'ERR',repr(ex)[:200], flush=True)
