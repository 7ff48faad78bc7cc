"""Context (sequence) parallel attention (reference torchacc/ops/context_parallel/__init__.py:1-7)."""
from .comm import (all_gather, diff_all_to_all, gather_forward_split_backward, seq_head_all_to_all,
                   split_forward_gather_backward)
from .groups import (get_context_parallel_group, get_context_parallel_size, get_inter_cp_process_group,
                     get_intra_cp_process_group, initialize_context_parallel, initialize_parallel_group, use_mesh)
from .ring import merge_out_lse, ring_attention, zigzag_positions, zigzag_split
from .two_d import context_parallel_2d
from .ulysses import ulysses
from .model_hook import cp_attention_qkvpacked, cp_local_positions, cp_shard_sequence

__all__ = ["ulysses", "ring_attention", "context_parallel_2d", "initialize_context_parallel",
           "initialize_parallel_group", "get_context_parallel_group", "get_inter_cp_process_group",
           "get_intra_cp_process_group", "get_context_parallel_size", "split_forward_gather_backward",
           "gather_forward_split_backward", "seq_head_all_to_all", "diff_all_to_all", "all_gather", "merge_out_lse",
           "zigzag_split", "zigzag_positions", "use_mesh", "cp_attention_qkvpacked", "cp_local_positions",
           "cp_shard_sequence"]
