#!/usr/bin/env python
"""A handful of single launches of the round-2 kernels for `ncu --set full` (one GPU; see tools/gpu_validate.sh):
MX-FP8 GEMM, gate|up GEMM with the SwiGLU epilogue, blockwise attention forward with the in-epilogue merge, attention
forward with dropout.  Shapes are the Llama-3-8B ones at T = 8192 tokens."""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from torchacc_b200.ops import attention as A  # noqa: E402
from torchacc_b200.ops import fp8  # noqa: E402
from torchacc_b200.ops.context_parallel import ring as R  # noqa: E402
from torchacc_b200.ops.swiglu import gate_up_swiglu  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    T = 8192
    x = (torch.randn(T, 4096, device=dev) * 0.5).bfloat16()
    w_qkv = (torch.randn(6144, 4096, device=dev) * 0.05).bfloat16()
    w_gu = (torch.randn(28672, 4096, device=dev) * 0.05).bfloat16()
    for _ in range(2):                                       # second launch = warm instruction cache / TMA descriptors
        xq, wq = fp8.quantize_mxfp8(x)[0], fp8.quantize_mxfp8(w_qkv)[0]
        fp8.gemm_mxfp8(xq, wq)
        with torch.no_grad():
            gate_up_swiglu(x, w_gu)
        B, S, Hq, Hk, D, cp = 2, 1024, 32, 8, 128, 4
        q = (torch.randn(B, S, Hq, D, device=dev) * 0.8).bfloat16()
        blocks = [(torch.randn(2, B, S, Hk, D, device=dev) * 0.8).bfloat16() for _ in range(cp)]
        R.ring_forward_native(q, blocks, R._plan(cp - 1, cp, True, True), 1.0 / math.sqrt(D))
        k, v = blocks[0][0], blocks[0][1]
        A.flash_attn_func(q, k, v, dropout_p=0.1, causal=True)
    torch.cuda.synchronize()
    print("ncu targets done")


if __name__ == "__main__":
    main()
