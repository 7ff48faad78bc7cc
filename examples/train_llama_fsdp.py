#!/usr/bin/env python
"""FSDP training of a native Llama model (the quick-start of docs/, counterpart of reference docs/source/dist/fsdp.md).

    torchrun --nproc-per-node 8 --master-addr 127.0.0.1 examples/train_llama_fsdp.py --model llama3-8b --seq_len 4096
    python examples/train_llama_fsdp.py --model tiny --seq_len 128 --steps 20          # single process / CPU
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "benchmarks"))
import torchacc_b200 as ta  # noqa: E402
from dataset import MarkovLM  # noqa: E402
from torchacc_b200.models import build_llama, llama_config  # noqa: E402


def main():
    p = argparse.ArgumentParser()
    p.add_argument("--model", default="tiny")
    p.add_argument("--seq_len", type=int, default=128)
    p.add_argument("--batch_size", type=int, default=2)
    p.add_argument("--steps", type=int, default=50)
    p.add_argument("--ckpt_dir", default=None, help="save a sharded checkpoint here at the end")
    a = p.parse_args()

    device = ta.dist.current_device()
    bf16 = device.type == "cuda"
    mcfg = llama_config(a.model, vocab_size=4096) if a.model == "tiny" else llama_config(a.model)
    with torch.device(device):
        model = build_llama(mcfg, dtype=torch.bfloat16 if bf16 else torch.float32)

    cfg = ta.Config()
    cfg.compute.bf16 = bf16
    cfg.memory.gc = True                                   # recompute each decoder layer in backward
    cfg.dist.fsdp.size = ta.dist.world_size()              # ZeRO-3 over all ranks
    cfg.dist.fsdp.wrap_layer_cls = {"LlamaDecoderLayer"}
    cfg.dataloader.prefetch = 2

    ds = MarkovLM(mcfg.vocab_size, a.seq_len, num_samples=1024)
    sampler = torch.utils.data.distributed.DistributedSampler(ds, ta.dist.world_size(), ta.dist.rank()) \
        if ta.dist.world_size() > 1 else None
    loader = torch.utils.data.DataLoader(ds, batch_size=a.batch_size, sampler=sampler, shuffle=sampler is None,
                                         drop_last=True)
    model, loader = ta.accelerate(model, loader, cfg)       # AsyncLoader: pinned H2D copies one batch ahead
    opt = ta.optim.FusedAdamW(model.parameters(), lr=3e-4, betas=(0.9, 0.95), weight_decay=0.1)

    step = 0
    while step < a.steps:
        for batch in loader:
            loss = model(**batch)["loss"]
            loss.backward()
            model.clip_grad_norm_(1.0)
            opt.step()
            model.zero_grad()
            if step % 10 == 0 and ta.dist.rank() == 0:
                print(f"step {step:4d}  loss {float(loss):.4f}  (optimum {ds.optimal_loss:.3f})", flush=True)
            step += 1
            if step >= a.steps:
                break

    if a.ckpt_dir:
        # sharded model + optimizer state; `python -m torchacc_b200.utils.consolidate_and_reshard_ckpts` converts it
        from torchacc_b200.parallel.state_dict_utils import save_sharded_checkpoint
        save_sharded_checkpoint(model, opt, a.ckpt_dir)


if __name__ == "__main__":
    main()
