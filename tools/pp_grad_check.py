#!/usr/bin/env python
"""Diagnostic (2 GPUs): per-parameter gradients of one pipeline-parallel step (pp=2, M micro-batches) against the same
batch on one GPU.  torchrun --nproc-per-node 2 tools/pp_grad_check.py [--mb 2] [--no-clip]"""
import argparse
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torchacc_b200 as ta  # noqa: E402
from torchacc_b200.models import LlamaDecoderLayer, build_llama  # noqa: E402
from torchacc_b200.parallel.fsdp import ShardingEngine, shard_model  # noqa: E402


def tiny(dev):
    torch.manual_seed(0)
    with torch.device(dev):
        return build_llama("tiny", hidden_size=512, intermediate_size=1024, num_hidden_layers=2, num_attention_heads=8,
                           num_key_value_heads=2, head_dim=64, vocab_size=2048, max_position_embeddings=512,
                           dtype=torch.bfloat16)


def grads_of(engine):
    out = {}
    for u in engine.units:
        g = getattr(u.flat_param, "_tb_grad", None)
        if g is None:
            g = u.flat_param.grad
        if g is None:
            continue
        for info in u.infos:
            name = (u.prefix + "." if u.prefix else "") + info.fqn
            out[name] = g[info.offset:info.offset + info.numel].float().clone()
    return out


def canon(name, layer_offset=0):
    name = name.replace("model.", "")
    parts = name.split(".")
    if parts[0] == "layers":
        parts[1] = str(int(parts[1]) + layer_offset)
    return ".".join(parts)


def main():
    p = argparse.ArgumentParser()
    p.add_argument("--mb", type=int, default=2)
    a = p.parse_args()
    rank = int(os.environ["LOCAL_RANK"])
    dev = torch.device("cuda", rank)
    torch.cuda.set_device(dev)
    g = torch.Generator().manual_seed(1)
    ids = torch.randint(0, 2048, (4, 128), generator=g).to(dev)
    # single GPU
    ref_model = tiny(dev)
    eng = ShardingEngine(dev, compute_dtype=torch.bfloat16, strategy="NO_SHARD", grad_mode="compat")
    root = shard_model(ref_model, eng, (LlamaDecoderLayer,), ())
    opt = ta.optim.FusedAdamW(eng.flat_parameters(), lr=1e-3)
    ref_hidden = []
    hk = ref_model.model.layers[0].register_forward_hook(lambda m, i, o: ref_hidden.append(o.detach().float().clone()))
    for part in ids.chunk(a.mb):        # per-micro-batch activations after layer 0 (no grads kept)
        with torch.no_grad():
            root(input_ids=part, labels=part)
    hk.remove()
    ref_hidden = ref_hidden[:a.mb]
    loss = root(input_ids=ids, labels=ids)["loss"]
    loss.backward()
    ref = {canon(k): v for k, v in grads_of(eng).items()}
    ref_norm = float(eng.clip_grad_norm_(1.0))
    # pipeline
    model = tiny(dev)
    cfg = ta.Config()
    cfg.compute.bf16 = True
    cfg.dist.pp.size = 2
    cfg.dist.pp.num_micro_batches = a.mb
    cfg.dist.pp.split_points = ["model.layers.1"]
    model = ta.accelerate(model, config=cfg)
    opt2 = ta.optim.FusedAdamW(model.parameters(), lr=1e-3)
    seen = []
    stage_mod = model.pp_wrapper.executor.module
    inner = stage_mod
    while hasattr(inner, "module") and not hasattr(inner, "layers"):
        inner = inner.module
    if rank == 1:
        seen_kw = []

        def _hook(m, args, kwargs):
            if "hidden" in kwargs and len(seen_kw) < 4:
                seen.append(kwargs["hidden"].detach().float().clone())
                seen_kw.append({k: (v.detach().clone() if isinstance(v, torch.Tensor) else v) for k, v in kwargs.items()})
                print("  stage-1 call kwargs:", {k: (tuple(v.shape), str(v.dtype), v.stride()) if isinstance(v, torch.Tensor)
                                                  else v for k, v in kwargs.items()}, "grad", torch.is_grad_enabled(),
                      "autocast", torch.is_autocast_enabled(), flush=True)
        inner.register_forward_pre_hook(_hook, with_kwargs=True)
        inner.register_forward_hook(lambda m, a, o: print("  stage-1 forward returned loss", float(o["loss"]), flush=True))
    pl = model.forward_backward(input_ids=ids, labels=ids, output_fn=lambda out: out["loss"])
    if rank == 1:
        for i, h in enumerate(seen):
            r = ref_hidden[i].reshape(h.shape)
            print(f"  stage-1 input mb{i}: |h| {float(h.norm()):.4f} ref {float(r.norm()):.4f} relerr "
                  f"{float((h - r).norm() / r.norm()):.3e}", flush=True)
    if rank == 1:
        # weights of this stage vs the single-GPU model, and both halves run directly on the captured input
        ref_l1 = ref_model.model.layers[1]
        ref_l1 = getattr(ref_l1, "module", ref_l1)
        for (n1, p1), (n2, p2) in zip([(i.fqn, i.tensor) for i in model.engine.units[-1].infos],
                                       [(i.fqn, i.tensor) for u in eng.units for i in u.infos
                                        if (u.prefix + "." + i.fqn).startswith("model.layers.1") or u.prefix == ""
                                        and ("norm" in i.fqn or "lm_head" in i.fqn)]):
            print(f"  weight {n1:42s} vs {n2:42s} maxdiff {float((p1.float() - p2.float()).abs().max()):.3e}", flush=True)
        with torch.no_grad():
            hid = seen[0]
            lab_ids = ids.chunk(a.mb)[0]
            out_stage = inner(hidden=hid.to(torch.bfloat16), labels=lab_ids)["loss"]
            core = ref_model.model
            core = getattr(core, "module", core)
            B, S = hid.shape[0], hid.shape[1]
            h2 = ref_l1(hid.to(torch.bfloat16).reshape(B * S, -1), core.rope(hid.device), B, S, None, None, core.pctx)
            from torchacc_b200.ops import fused_linear_cross_entropy, rmsnorm
            y2, _ = rmsnorm(h2, core.norm.weight, core.norm.eps)
            lab = torch.full_like(lab_ids, -100)
            lab[:, :-1] = lab_ids[:, 1:]
            l2 = fused_linear_cross_entropy(y2, ref_model.lm_head.weight, lab.reshape(-1))
            print(f"  direct: stage module loss {float(out_stage):.5f}  reference second half {float(l2):.5f}", flush=True)
        kw0 = seen_kw[0]
        print("  labels equal ids chunk:", bool(torch.equal(kw0["labels"], lab_ids)), flush=True)
        with torch.no_grad():
            print("  replay of the captured kwargs:", float(inner(**kw0)["loss"]), flush=True)
        hb = hid.to(torch.bfloat16)
        la = float(inner(hidden=hb.clone().requires_grad_(), labels=lab_ids)["loss"])
        with torch.autocast("cuda", dtype=torch.bfloat16):
            lb = float(inner(hidden=hb.clone().requires_grad_(), labels=lab_ids)["loss"])
        lc = float(model.pp_wrapper.executor.module(hidden=hb.clone().requires_grad_(), labels=lab_ids)["loss"])
        with torch.no_grad():
            ld = float(model.pp_wrapper.executor.module(hidden=hb.clone(), labels=lab_ids)["loss"])
        r1, r2 = inner._owner[0].rope(hid.device), core.rope(hid.device)
        print(f"  grad-mode {la:.5f}  grad+autocast {lb:.5f}  through engine {lc:.5f}  engine no_grad {ld:.5f}  "
              f"rope diff {float((r1[0] - r2[0]).abs().max()):.3e}", flush=True)
    got = {canon(k, layer_offset=rank): v for k, v in grads_of(model.engine).items()}
    norm = float(model.clip_grad_norm_(1.0))
    torch.cuda.synchronize()
    for r in range(2):
        dist.barrier()
        if r != rank:
            continue
        print(f"--- rank {rank}: loss single {float(loss):.5f} pp {float(pl) if pl is not None else float('nan'):.5f}  "
              f"grad norm single {ref_norm:.4f} pp {norm:.4f}", flush=True)
        for k, v in got.items():
            if k not in ref:
                print(f"  {k}: no reference entry ({list(ref)[:3]}...)")
                continue
            rv = ref[k]
            rel = float((v - rv).norm() / (rv.norm() + 1e-12))
            ratio = float(v.norm() / (rv.norm() + 1e-12))
            print(f"  {k:45s} |g| {float(v.norm()):.4e} ref {float(rv.norm()):.4e} ratio {ratio:.3f} relerr {rel:.3e}", flush=True)
    dist.barrier()


if __name__ == "__main__":
    main()
