"""Per-phase timeline of one backward-attention CTA (clock64 stamps written by the kernel itself)."""
import ctypes, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from torchacc_b200 import _native as nat
from torchacc_b200.ops import attention as A
A.set_attention_backend("native")
L = nat.require()
L.tb_flash_attn_bwd_set_trace.argtypes = [ctypes.c_uint64]
B, S, Hq, Hk, D = 2, 4096, 32, 8, 128
q = torch.randn(B, S, Hq, D, device="cuda", dtype=torch.bfloat16, requires_grad=True)
k = torch.randn(B, S, Hk, D, device="cuda", dtype=torch.bfloat16, requires_grad=True)
v = torch.randn(B, S, Hk, D, device="cuda", dtype=torch.bfloat16, requires_grad=True)
do = torch.randn(B, S, Hq, D, device="cuda", dtype=torch.bfloat16)
o = A.flash_attn_func(q, k, v, causal=True); o.backward(do)
trace = torch.zeros(64 * 16, dtype=torch.int64, device="cuda")
L.tb_flash_attn_bwd_set_trace(trace.data_ptr())
o = A.flash_attn_func(q, k, v, causal=True); o.backward(do)
torch.cuda.synchronize()
L.tb_flash_attn_bwd_set_trace(0)
t = trace.view(64, 16).cpu()
names = ["mma:qdo_full", "mma:r1_free", "mma:ab_issued", "mma:pds_ready", "mma:edc_issued", "sm:sdp_full", "sm:S_read+bar",
         "sm:pds_arrive", "sm:dq_full", "sm:r1_arrive", "sm:ph1_done", "sm:stage_free", "sm:dp_full", "sm:ph2_done"]
t0 = int(t[2, 0])
for it in range(2, 10):
    row = [(names[j], int(t[it, j]) - t0) for j in range(len(names))]
    print(it, " ".join(f"{n}={c}" for n, c in sorted(row, key=lambda x: x[1])))
per = (int(t[40, 0]) - int(t[8, 0])) / 32
print("cycles per iteration (steady):", per)

# ---- forward kernel: first CTA = the LAST (heaviest) causal query tile of head 0 / batch 0 ----
L.tb_flash_attn_fwd_set_trace.argtypes = [ctypes.c_uint64]
ftrace = torch.zeros(64 * 16, dtype=torch.int64, device="cuda")
L.tb_flash_attn_fwd_set_trace(ftrace.data_ptr())
o = A.flash_attn_func(q, k, v, causal=True)
torch.cuda.synchronize()
L.tb_flash_attn_fwd_set_trace(0)
f = ftrace.view(64, 16).cpu()
fn = ["mma:S(t+1)_issued", "mma:p_ready", "sm:s_full", "sm:max_xchg", "sm:exp_done", "sm:o_done", "sm:p_arrive"]
f0 = int(f[2, 2])
for tt in range(2, 8):
    row = [(fn[j], int(f[tt, j]) - f0) for j in range(len(fn))]
    print("fwd", tt, " ".join(f"{n}={c}" for n, c in sorted(row, key=lambda x: x[1])))
print("fwd cycles per tile (steady):", (int(f[28, 6]) - int(f[4, 6])) / 24)
