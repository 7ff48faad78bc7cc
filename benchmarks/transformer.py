#!/usr/bin/env python
"""Causal-LM training benchmark (counterpart of reference benchmarks/transformer.py:32-207).

Same knobs as the reference script (``--dp_size --tp_size --pp_size --fsdp_size --sp_size --gc --bf16 --fp16
--profile --log_interval``), same default model (GPT-2, seq 512, batch 8) and metric (samples/s), with two
differences: data is synthetic (no network for wikitext), and throughput is timed on the DEVICE with CUDA events and
reduced with MAX over ranks (the reference uses rank-0 host wall-clock between log points).

    torchrun --nproc-per-node 4 --master-addr 127.0.0.1 benchmarks/transformer.py --model_name llama3.2-1b \\
             --fsdp_size 4 --gc --bf16 --max_seq_length 2048
"""
import argparse
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torchacc_b200 as ta  # noqa: E402
from torchacc_b200.models import build_gpt2, build_llama  # noqa: E402
from torchacc_b200.models.gpt2 import PRESETS as GPT2_PRESETS  # noqa: E402


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--model_name", default="gpt2")
    p.add_argument("--max_seq_length", type=int, default=512)
    p.add_argument("--batch_size", type=int, default=8, help="per-rank batch size")
    p.add_argument("--num_train_steps", type=int, default=30)
    p.add_argument("--log_interval", type=int, default=10)
    p.add_argument("--lr", type=float, default=1e-4)
    p.add_argument("--dp_size", type=int, default=None)
    p.add_argument("--tp_size", type=int, default=1)
    p.add_argument("--pp_size", type=int, default=1)
    p.add_argument("--pp_micro_batches", type=int, default=4)
    p.add_argument("--fsdp_size", type=int, default=1)
    p.add_argument("--sp_size", type=int, default=1)
    p.add_argument("--sp_mode", default="ulysses")
    p.add_argument("--gc", action="store_true")
    p.add_argument("--bf16", action="store_true")
    p.add_argument("--fp16", action="store_true")
    p.add_argument("--profile", action="store_true", help="torch.profiler trace to ./log/profile")
    p.add_argument("--disable_loss_print", action="store_true")
    p.add_argument("--tp_no_sequence_parallel", action="store_true")
    p.add_argument("--hf", action="store_true",
                   help="build the HuggingFace LlamaForCausalLM of the preset's geometry instead of the native model "
                        "(the reference benchmark uses AutoModelForCausalLM.from_config, benchmarks/transformer.py:96-102); "
                        "tp / sp / pp / fsdp apply to it through accelerate()")
    return p.parse_args()


def main():
    a = parse()
    rank, world = ta.dist.rank(), ta.dist.world_size()
    device = ta.dist.current_device()
    if device.type == "cuda":
        torch.cuda.set_device(device)
    dtype = torch.bfloat16 if a.bf16 else (torch.float16 if a.fp16 else torch.float32)
    is_gpt2 = a.model_name in GPT2_PRESETS
    with torch.device(device):
        if is_gpt2:
            model = build_gpt2(a.model_name, dtype=dtype)
        else:
            from torchacc_b200.models import llama_config
            base = llama_config(a.model_name)
            max_pos = max(a.max_seq_length, base.max_position_embeddings)
            if a.hf:
                from transformers import LlamaConfig, LlamaForCausalLM
                hc = LlamaConfig(vocab_size=base.vocab_size, hidden_size=base.hidden_size,
                                 intermediate_size=base.intermediate_size, num_hidden_layers=base.num_hidden_layers,
                                 num_attention_heads=base.num_attention_heads,
                                 num_key_value_heads=base.num_key_value_heads, head_dim=base.head_dim,
                                 max_position_embeddings=max_pos, rope_theta=base.rope_theta,
                                 rms_norm_eps=base.rms_norm_eps, tie_word_embeddings=False, use_cache=False,
                                 attn_implementation="flash_attention_2" if device.type == "cuda" else "eager")
                model = LlamaForCausalLM(hc).to(dtype)
            else:
                model = build_llama(a.model_name, dtype=dtype, max_position_embeddings=max_pos)
    if a.hf and is_gpt2:
        raise SystemExit("--hf builds Llama-family presets")
    layer_cls = "GPT2Block" if is_gpt2 else "LlamaDecoderLayer"
    vocab = model.config.vocab_size

    cfg = ta.Config()
    cfg.compute.bf16, cfg.compute.fp16 = a.bf16, a.fp16
    cfg.memory.gc = a.gc
    cfg.memory.gc_cls = {layer_cls}
    cfg.dist.dp.size = a.dp_size
    cfg.dist.tp.size, cfg.dist.fsdp.size, cfg.dist.sp.size = a.tp_size, a.fsdp_size, a.sp_size
    cfg.dist.sp.mode = a.sp_mode
    if a.tp_no_sequence_parallel:
        cfg.dist.tp.sequence_parallel = False
    cfg.dist.fsdp.wrap_layer_cls = {layer_cls}
    if a.pp_size > 1:
        n_layers = model.config.n_layer if is_gpt2 else model.config.num_hidden_layers
        per = n_layers // a.pp_size
        prefix = "h" if is_gpt2 else "model.layers"
        cfg.dist.pp.size = a.pp_size
        cfg.dist.pp.num_micro_batches = a.pp_micro_batches
        cfg.dist.pp.split_points = [f"{prefix}.{per * i}" for i in range(1, a.pp_size)]

    # ranks that differ only in their pp / sp / tp coordinate work on the SAME samples (default topology: those axes vary
    # fastest), so the data stream is seeded by the data-parallel index
    g = torch.Generator().manual_seed(rank // (a.pp_size * a.sp_size * a.tp_size))
    batches = [{"input_ids": torch.randint(0, vocab, (a.batch_size, a.max_seq_length), generator=g)} for _ in range(8)]
    for b in batches:
        b["labels"] = b["input_ids"]

    class Loader:
        def __len__(self):
            return a.num_train_steps

        def __iter__(self):
            for i in range(a.num_train_steps):
                yield batches[i % len(batches)]

    model, loader = ta.accelerate(model, Loader(), cfg)
    opt = ta.optim.FusedAdamW(model.parameters(), lr=a.lr)
    scaler = ta.amp.GradScaler(enabled=a.fp16)
    prof = None
    if a.profile:
        from torch.profiler import ProfilerActivity, profile, schedule, tensorboard_trace_handler
        prof = profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], schedule=schedule(wait=2, warmup=2, active=6),
                       on_trace_ready=tensorboard_trace_handler("./log/profile"))
        prof.start()

    def hf_loss(logits, labels):          # last pipeline stage of an HF model returns the logits
        return torch.nn.functional.cross_entropy(logits[:, :-1].reshape(-1, logits.shape[-1]).float(),
                                                 labels[:, 1:].reshape(-1))

    cuda = device.type == "cuda"
    t_ev = torch.cuda.Event(enable_timing=True) if cuda else None
    if cuda:
        t_ev.record()
    t_host = time.perf_counter()
    for step, batch in enumerate(loader, 1):
        if a.pp_size > 1:
            loss = model.forward_backward(**batch, output_fn=hf_loss) if a.hf else model.forward_backward(**batch)
        else:
            out = model(**batch)
            loss = out["loss"] if isinstance(out, dict) else out.loss
            scaler.scale(loss).backward()
        model.clip_grad_norm_(1.0)
        scaler.step(opt)
        scaler.update()
        model.zero_grad()
        if prof is not None:
            prof.step()
        if step % a.log_interval == 0:
            if cuda:
                e = torch.cuda.Event(enable_timing=True)
                e.record()
                e.synchronize()
                dt = t_ev.elapsed_time(e) / 1e3
                t_ev = e
            else:
                now = time.perf_counter()
                dt, t_host = now - t_host, now
            t = torch.tensor([dt], device=device)
            if world > 1:
                torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
            replicas = world // (a.tp_size * a.pp_size * a.sp_size)
            sps = a.batch_size * replicas * a.log_interval / float(t)
            if rank == 0:
                rec = {"step": step, "samples_per_s": round(sps, 2), "tokens_per_s": round(sps * a.max_seq_length, 1)}
                if not a.disable_loss_print:
                    rec["loss"] = round(float(loss), 4)
                print(json.dumps(rec), flush=True)
    if prof is not None:
        prof.stop()


if __name__ == "__main__":
    main()
