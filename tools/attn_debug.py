import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from unet_design_b200._lib import ops as raw
o_ = raw()
torch.manual_seed(0)
R, C, T = 256, 64, 64
qkv = torch.randn(R, 3 * C, device="cuda").to(torch.bfloat16)
q, k, v = qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:]
P = torch.empty(R, 256, dtype=torch.bfloat16, device="cuda")
o_.bgemm256(q, False, k, False, P, C, 1, C ** -0.5 * 1.4426950408889634, T, None)
torch.cuda.synchronize(); print("fwd1 ok", float(P.float().sum(1).mean()))
out = torch.empty(R, C, dtype=torch.bfloat16, device="cuda")
o_.bgemm256(P, False, v, True, out, 256, 0, 1.0, T, None)
torch.cuda.synchronize(); print("fwd2 ok")
go = torch.randn(R, C, device="cuda").to(torch.bfloat16)
dS = torch.empty(R, 256, dtype=torch.bfloat16, device="cuda")
for name, args in (("dS", (go, False, v, False, dS, C, 2, C ** -0.5, T, P)),):
    print(name, [(tuple(t.shape), t.stride(), t.data_ptr() % 256) for t in args if torch.is_tensor(t)])
    try:
        o_.bgemm256(*args); torch.cuda.synchronize(); print(name, "ok")
    except Exception as e:
        print(name, "FAILED", str(e)[:200])
dqkv = torch.empty(R, 3 * C, dtype=torch.bfloat16, device="cuda")
for name, args in (("dV", (P, True, go, True, dqkv[:, 2 * C:], 256, 0, 1.0, T, None)), ("dQ", (dS, False, k, True, dqkv[:, :C], 256, 0, 1.0, T, None)),
                   ("dK", (dS, True, q, True, dqkv[:, C:2 * C], 256, 0, 1.0, T, None))):
    try:
        o_.bgemm256(*args); torch.cuda.synchronize(); print(name, "ok")
    except Exception as e:
        print(name, "FAILED", str(e)[:200])

print("---- through autograd")
from unet_design_b200 import ops
import unet_design_b200.ops as O
orig = O._AttnCore.backward
def dbg(ctx, go):
    print("go", tuple(go.shape), go.stride(), go.dtype, go.data_ptr() % 256, go.is_contiguous())
    return orig(ctx, go)
O._AttnCore.backward = staticmethod(dbg)
n, tokens, c = 4, 64, 64
qkv4 = (torch.randn(n, 8, 8, 3 * c, device="cuda") * 1.5).to(torch.bfloat16).requires_grad_(True)
o = ops.attention_core(qkv4)
g = torch.randn_like(o)
try:
    o.backward(g); torch.cuda.synchronize(); print("autograd ok", float(qkv4.grad.float().abs().mean()))
except Exception as e:
    print("autograd FAILED", str(e)[:300])
