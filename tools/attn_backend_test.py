import torch, time
import torch.nn.functional as F
from torch.nn.attention import sdpa_kernel, SDPBackend
n,s,c=128,256,256
qkv=torch.randn(n,1,s,3*c,device='cuda',dtype=torch.bfloat16,requires_grad=True)
def run(backend):
    q,k,v=qkv[...,:c],qkv[...,c:2*c],qkv[...,2*c:]
    with sdpa_kernel(backend):
        o=F.scaled_dot_product_attention(q,k,v,scale=c**-0.5)
    g=torch.randn_like(o)
    def fb():
        o=F.scaled_dot_product_attention(q,k,v,scale=c**-0.5)
        o.backward(g)
    with sdpa_kernel(backend):
        for _ in range(3): fb()
        torch.cuda.synchronize()
        e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): fb()
        e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/20*1e3
for b in [SDPBackend.FLASH_ATTENTION, SDPBackend.CUDNN_ATTENTION, SDPBackend.EFFICIENT_ATTENTION]:
    try:
        print(b, '%.1f us fwd+bwd'%run(b))
    except Exception as e:
        print(b,'failed',str(e)[:200])
