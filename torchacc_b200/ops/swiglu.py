"""SwiGLU: ``h = silu(g) * u`` on the kernels in csrc/ops/norm_rope_act.cu.

``swiglu(gu)`` takes the fused gate|up projection output ``[..., 2F]``; ``swiglu_separate(g, u)`` takes the two
halves as separate (row-strided) tensors, which is what the HF module patches use.
Replaces liger's Triton SiLU-mul (reference torchacc/ops/liger.py:21-28)."""
from __future__ import annotations

import torch
import torch.nn.functional as F

from .. import _native as nat


def _rows(t):
    t2 = t.reshape(-1, t.shape[-1])
    if t2.stride(-1) != 1 or t2.stride(0) % 8 != 0:
        t2 = t2.contiguous()
    return t2


class _SwiGLUFn(torch.autograd.Function):

    @staticmethod
    def forward(ctx, g, u):
        Fdim = g.shape[-1]
        g2, u2 = _rows(g), _rows(u)
        T = g2.shape[0]
        h = torch.empty((*g.shape[:-1], Fdim), dtype=g.dtype, device=g.device)
        L = nat.require()
        nat.check(L.tb_swiglu_fwd(g2.data_ptr(), u2.data_ptr(), h.data_ptr(), T, Fdim, g2.stride(0), u2.stride(0),
                                  nat.num_sms(), nat.stream(), nat.bf16_flag(g2)), "tb_swiglu_fwd")
        nat.count_launch()
        ctx.save_for_backward(g2, u2)
        ctx.shape = g.shape
        return h

    @staticmethod
    def backward(ctx, dh):
        g2, u2 = ctx.saved_tensors
        T, Fdim = g2.shape
        dh2 = dh.contiguous().view(T, Fdim)
        dgu = torch.empty((T, 2 * Fdim), dtype=g2.dtype, device=g2.device)
        dg, du = dgu[:, :Fdim], dgu[:, Fdim:]
        L = nat.require()
        nat.check(
            L.tb_swiglu_bwd(dh2.data_ptr(), g2.data_ptr(), u2.data_ptr(), dg.data_ptr(), du.data_ptr(), T, Fdim,
                            g2.stride(0), u2.stride(0), dgu.stride(0), dgu.stride(0), nat.num_sms(), nat.stream(),
                            nat.bf16_flag(g2)),
            "tb_swiglu_bwd")
        nat.count_launch()
        return dg.view(ctx.shape), du.view(ctx.shape)


def swiglu_separate(g: torch.Tensor, u: torch.Tensor) -> torch.Tensor:
    if nat.is_half(g, u) and nat.use_native(g, u) and g.shape[-1] % 8 == 0:
        return _SwiGLUFn.apply(g, u)
    return (F.silu(g.float()) * u.float()).to(g.dtype)


class _SwiGLUFusedFn(torch.autograd.Function):
    """Fused gate|up input: the gradient is produced as ONE [T, 2F] tensor (no slice-gradient scatter)."""

    @staticmethod
    def forward(ctx, gu):
        Fdim = gu.shape[-1] // 2
        gu2 = _rows(gu)
        T = gu2.shape[0]
        h = torch.empty((*gu.shape[:-1], Fdim), dtype=gu.dtype, device=gu.device)
        L = nat.require()
        nat.check(L.tb_swiglu_fwd(gu2.data_ptr(), gu2.data_ptr() + 2 * Fdim, h.data_ptr(), T, Fdim, gu2.stride(0),
                                  gu2.stride(0), nat.num_sms(), nat.stream(), nat.bf16_flag(gu2)), "tb_swiglu_fwd")
        nat.count_launch()
        ctx.save_for_backward(gu2)
        ctx.shape = gu.shape
        return h

    @staticmethod
    def backward(ctx, dh):
        (gu2,) = ctx.saved_tensors
        T, Fdim = gu2.shape[0], gu2.shape[1] // 2
        dh2 = dh.contiguous().view(T, Fdim)
        dgu = torch.empty(ctx.shape, dtype=gu2.dtype, device=gu2.device)
        ld = 2 * Fdim
        L = nat.require()
        nat.check(
            L.tb_swiglu_bwd(dh2.data_ptr(), gu2.data_ptr(), gu2.data_ptr() + 2 * Fdim, dgu.data_ptr(),
                            dgu.data_ptr() + 2 * Fdim, T, Fdim, gu2.stride(0), gu2.stride(0), ld, ld, nat.num_sms(),
                            nat.stream(), nat.bf16_flag(gu2)), "tb_swiglu_bwd")
        nat.count_launch()
        return dgu


def swiglu(gu: torch.Tensor) -> torch.Tensor:
    """gu: [..., 2F] with gate in the first half and up in the second half."""
    Fdim = gu.shape[-1] // 2
    if nat.is_half(gu) and nat.use_native(gu) and Fdim % 8 == 0:
        return _SwiGLUFusedFn.apply(gu)
    return (F.silu(gu[..., :Fdim].float()) * gu[..., Fdim:].float()).to(gu.dtype)
