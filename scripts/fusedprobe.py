import sys, torch
sys.path.insert(0, 'alphazero-general_b200')
from azb200 import nnet as aznet
from azb200.fused_nn import FusedResNetEvaluator
dev = torch.device('cuda')
torch.manual_seed(0)
m = aznet.ResNet((4, 6, 7), 7, 3, **aznet.DEFAULT_NET_ARGS).to(dev).eval()
for B in (4096, 8192, 16384):
    obs = torch.rand(B, 4, 6, 7, device=dev); pol = torch.empty(B, 7, device=dev); val = torch.empty(B, 3, device=dev)
    ev = FusedResNetEvaluator(m, obs, pol, val)
    for _ in range(5): ev()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50): ev()
    e1.record(); torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / 50
    flops = B * (42 * 9 * 16 * 32 * 2 + 8 * 42 * 288 * 32 * 2)
    print(f'fused B={B}: {t*1000:.0f} us -> {B/t/1e3:.2f} M evals/s, {flops/t/1e9:.1f} TFLOP/s (MMA flops issued)', flush=True)
