#!/usr/bin/env python
"""An unmodified HuggingFace causal LM through ``ta.accelerate`` (counterpart of the reference's HF tutorial and
``examples/train_olmo.ipynb``): class-level kernel patches (RMSNorm / SwiGLU / fused linear-CE on our kernels), FSDP by
decoder-layer class name, activation checkpointing, bf16.

    torchrun --nproc-per-node 2 --master-addr 127.0.0.1 examples/train_hf_model.py
    torchrun --nproc-per-node 4 --master-addr 127.0.0.1 examples/train_hf_model.py --tp 2      # tp2 x fsdp2
    torchrun --nproc-per-node 4 --master-addr 127.0.0.1 examples/train_hf_model.py --sp 2      # Ulysses cp2 x fsdp2

The same HF object also runs under tensor parallelism (projections sharded in place) and context parallelism (sequence
sharded, attention over the sp group); ranks of one tp / sp group consume the same batch.
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torchacc_b200 as ta  # noqa: E402
from transformers import LlamaConfig, LlamaForCausalLM  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tp", type=int, default=1)
    ap.add_argument("--sp", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    a = ap.parse_args()
    device = ta.dist.current_device()
    hf_cfg = LlamaConfig(vocab_size=8192, hidden_size=512, intermediate_size=1408, num_hidden_layers=4,
                         num_attention_heads=8, num_key_value_heads=4, max_position_embeddings=1024,
                         attn_implementation="flash_attention_2" if device.type == "cuda" else "eager")
    model = LlamaForCausalLM(hf_cfg)                     # or LlamaForCausalLM.from_pretrained(<local path>)

    cfg = ta.Config()
    cfg.compute.bf16 = device.type == "cuda"
    cfg.memory.gc = True
    cfg.memory.gc_cls = {"LlamaDecoderLayer"}
    cfg.dist.tp.size, cfg.dist.sp.size = a.tp, a.sp
    cfg.dist.fsdp.size = ta.dist.world_size() // (a.tp * a.sp)
    cfg.dist.fsdp.wrap_layer_cls = {"LlamaDecoderLayer"}
    if device.type == "cuda":
        ta.utils.patch.patch_fa()                        # HF flash_attention_2 call sites -> our attention kernels
    model = ta.accelerate(model, config=cfg)
    opt = torch.optim.AdamW(model.parameters(), lr=1e-3)  # any torch optimizer works on the flat shards

    # data-parallel index: ranks that differ only in their tp / sp coordinate see the same samples
    g = torch.Generator().manual_seed(ta.dist.rank() // (a.tp * a.sp))
    ids = torch.randint(0, hf_cfg.vocab_size, (2, 256), generator=g).to(device)
    for step in range(a.steps):
        out = model(input_ids=ids, labels=ids)
        out.loss.backward()
        model.clip_grad_norm_(1.0)
        opt.step()
        model.zero_grad()
        if ta.dist.rank() == 0 and step % 5 == 0:
            print(f"step {step} loss {float(out.loss):.4f}")


if __name__ == "__main__":
    main()
