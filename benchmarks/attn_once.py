"""Run our attention fwd+bwd a few times (for ncu captures)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from torchacc_b200.ops import attention as A
A.set_attention_backend("native")
B, S, Hq, Hk, D = 2, 4096, 32, 8, 128
q = torch.randn(B, S, Hq, D, device="cuda", dtype=torch.bfloat16, requires_grad=True)
k = torch.randn(B, S, Hk, D, device="cuda", dtype=torch.bfloat16, requires_grad=True)
v = torch.randn(B, S, Hk, D, device="cuda", dtype=torch.bfloat16, requires_grad=True)
do = torch.randn(B, S, Hq, D, device="cuda", dtype=torch.bfloat16)
for _ in range(3):
    o = A.flash_attn_func(q, k, v, causal=True)
    o.backward(do)
torch.cuda.synchronize()
print("done")
