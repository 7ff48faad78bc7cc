"""CLI: consolidate or reshard FSDP checkpoints (reference torchacc/utils/consolidate_and_reshard_ckpts.py:12-157;
console script ``consolidate_and_reshard_fsdp_ckpts``).  Same flags as the reference; ``--ckpt_type model`` and
``--ckpt_type optimizer`` work on their own (they pass mismatching kwargs there)."""
from __future__ import annotations

import argparse


def build_parser() -> argparse.ArgumentParser:
    p = argparse.ArgumentParser(description="Consolidate and reshard sharded FSDP checkpoints (torchacc_b200 format)")
    p.add_argument("--ckpt_dir", type=str, required=True, help="directory that holds the rank*-of-* shard files")
    p.add_argument("--model_ckpt_name_pattern", type=str, default="rank*-of-*-model.pth")
    p.add_argument("--optimizer_ckpt_name_pattern", type=str, default="rank*-of-*-optim.pth")
    p.add_argument("--ckpt_type", type=str, default="all", choices=["all", "model", "optimizer"])
    p.add_argument("--reshard_num", type=int, default=1,
                   help="1 = write a single consolidated file; N > 1 = write N shards for an N-rank job")
    p.add_argument("--save_dir", type=str, default="", help="output directory (default: --ckpt_dir)")
    p.add_argument("--model_save_name_pattern", type=str, default="")
    p.add_argument("--optimizer_save_name_pattern", type=str, default="")
    return p


def main(argv=None) -> None:
    from ..parallel import state_dict_utils as U
    a = build_parser().parse_args(argv)
    if a.reshard_num < 1:
        raise SystemExit("--reshard_num must be >= 1")
    if a.ckpt_type in ("all", "model"):
        U.consolidate_and_reshard_fsdp_model_dict(a.ckpt_dir, a.model_ckpt_name_pattern, a.save_dir,
                                                  a.model_save_name_pattern, a.reshard_num)
    if a.ckpt_type in ("all", "optimizer"):
        U.consolidate_and_reshard_fsdp_optim_dict(a.ckpt_dir, a.optimizer_ckpt_name_pattern, a.save_dir,
                                                  a.optimizer_save_name_pattern, a.reshard_num)


if __name__ == "__main__":
    main()
