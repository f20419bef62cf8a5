"""GPU parity of the tcgen05 attention core (csrc/attention.cu: softmax(Q K^T / sqrt(C)) V as batched 256-row GEMMs with
fused softmax / softmax-backward epilogues) against a plain PyTorch fp32 reference of the same op on the same bf16 inputs
(diff_cifar/model.py:100-119).  Tolerance: bf16, <= 1e-2 norm-wise (north_star)."""
import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from unet_design_b200 import ops as o
    torch.backends.cuda.matmul.allow_tf32 = False
    return o


def _ref(qkv, tokens):
    n = qkv.shape[0] // tokens
    c = qkv.shape[1] // 3
    q, k, v = (t.reshape(n, tokens, c) for t in qkv.float().split(c, dim=1))
    w = torch.softmax(torch.bmm(q, k.transpose(1, 2)) * c ** -0.5, dim=-1)
    return torch.bmm(w, v).reshape(n * tokens, c)


@pytest.mark.parametrize("n,tokens,c", [(4, 64, 64), (2, 256, 128), (32, 16, 256), (8, 256, 256), (128, 256, 256), (16, 16, 192)])
def test_attention_core_forward_backward(ops, n, tokens, c):
    torch.manual_seed(0)
    assert ops.attention_core_supported(n * tokens, tokens, c)
    h = int(tokens ** 0.5)
    qkv = (torch.randn(n, h, tokens // h, 3 * c, device="cuda") * 1.5).to(torch.bfloat16).requires_grad_(True)
    o = ops.attention_core(qkv)
    g = torch.randn_like(o)
    o.backward(g)
    ref_in = qkv.detach().reshape(n * tokens, 3 * c).float().requires_grad_(True)
    ro = _ref(ref_in, tokens)
    ro.backward(g.reshape(n * tokens, c).float())
    assert o.shape == (n, h, tokens // h, c)
    assert rel_err(o.reshape(n * tokens, c), ro) < 1e-2, rel_err(o.reshape(n * tokens, c), ro)
    gq, gk, gv = qkv.grad.reshape(n * tokens, 3 * c).float().split(c, dim=1)
    rq, rk, rv = ref_in.grad.split(c, dim=1)
    assert rel_err(gv, rv) < 1e-2, ("dv", rel_err(gv, rv))
    assert rel_err(gq, rq) < 2e-2, ("dq", rel_err(gq, rq))
    assert rel_err(gk, rk) < 2e-2, ("dk", rel_err(gk, rk))


def test_attention_block_mask_keeps_samples_independent(ops):
    """Samples shorter than 256 tokens share a CTA: changing one sample must not change another one's output."""
    torch.manual_seed(1)
    qkv = torch.randn(16, 4, 4, 3 * 64, device="cuda").to(torch.bfloat16)
    o1 = ops.attention_core(qkv)
    qkv2 = qkv.clone()
    qkv2[3] = torch.randn_like(qkv2[3]) * 5
    o2 = ops.attention_core(qkv2)
    keep = [i for i in range(16) if i != 3]
    assert torch.equal(o1[keep], o2[keep]) and not torch.equal(o1[3], o2[3])
