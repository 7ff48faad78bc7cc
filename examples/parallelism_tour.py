#!/usr/bin/env python
"""One script, every parallel strategy: pick with --mode {dp,fsdp,hsdp,tp,pp,ulysses,ring,2d}.

    torchrun --nproc-per-node 4 --master-addr 127.0.0.1 examples/parallelism_tour.py --mode pp
(works on CPU with the gloo backend as well: `--nproc-per-node 2` on a laptop)
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torchacc_b200 as ta  # noqa: E402
from torchacc_b200.models import build_llama  # noqa: E402


def main():
    p = argparse.ArgumentParser()
    p.add_argument("--mode", default="fsdp")
    a = p.parse_args()
    world, rank = ta.dist.world_size(), ta.dist.rank()
    device = ta.dist.current_device()
    bf16 = device.type == "cuda"
    torch.manual_seed(0)
    with torch.device(device):
        model = build_llama("tiny", hidden_size=256, intermediate_size=512, num_hidden_layers=4,
                            num_attention_heads=8, num_key_value_heads=4, head_dim=32, vocab_size=1024,
                            max_position_embeddings=512, dtype=torch.bfloat16 if bf16 else torch.float32)
    cfg = ta.Config()
    cfg.compute.bf16 = bf16
    d = cfg.dist
    if a.mode == "dp":
        pass                                              # default: dp = world
    elif a.mode == "fsdp":
        d.fsdp.size = world
        d.fsdp.wrap_layer_cls = {"LlamaDecoderLayer"}
    elif a.mode == "hsdp":
        d.fsdp.size, d.dp.size = max(1, world // 2), world // max(1, world // 2)
        d.fsdp.wrap_layer_cls = {"LlamaDecoderLayer"}
    elif a.mode == "tp":
        d.tp.size = world
        d.tp.sequence_parallel = True
    elif a.mode == "pp":
        d.pp.size = world
        d.pp.num_micro_batches = 4
        d.pp.split_points = [f"model.layers.{i * 4 // world}" for i in range(1, world)]
    elif a.mode in ("ulysses", "ring", "2d"):
        d.sp.size = world
        d.sp.mode = a.mode
    model = ta.accelerate(model, config=cfg)
    opt = ta.optim.FusedAdamW(model.parameters(), lr=1e-3)
    data_rank = cfg.get_mesh().get_data_rank() if hasattr(cfg.get_mesh(), "get_data_rank") else rank
    ids = torch.randint(0, 1024, (4, 128), generator=torch.Generator().manual_seed(data_rank)).to(device)
    for step in range(10):
        if a.mode == "pp":
            loss = model.forward_backward(input_ids=ids, labels=ids, output_fn=lambda out: out["loss"])
        else:
            loss = model(input_ids=ids, labels=ids)["loss"]
            loss.backward()
        model.clip_grad_norm_(1.0)
        opt.step()
        model.zero_grad()
        if rank == world - 1 and loss is not None:
            print(f"[{a.mode}] step {step} loss {float(loss):.4f}", flush=True)


if __name__ == "__main__":
    main()
