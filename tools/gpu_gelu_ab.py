import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ln3diff_b200 import ops
dev="cuda"
torch.manual_seed(0)
M,K,N=12288,1024,4096
a=(torch.randn(M,K,device=dev)*0.5).bfloat16(); w=(torch.randn(N,K,device=dev)*0.03).bfloat16(); b=torch.randn(N,device=dev)
out=torch.empty(M,N,device=dev,dtype=torch.bfloat16)
def t(fn,n=30):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/n*1e3
us=t(lambda: ops.gemm(a,w,b,act=ops.ACT_GELU_ERF,out=out))
lin=(a.float()@w.float().t()+b)
ref=torch.nn.functional.gelu(lin)
rel=((out.float()-ref).norm()/ref.norm()).item()
mx=(out.float()-ref).abs().max().item()
print(f"exact={os.environ.get('LN3_GELU_EXACT','0')} fc1+gelu {us:.1f} us rel {rel:.3e} maxabs {mx:.3e}")
