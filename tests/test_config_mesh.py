import json
import os

import pytest
import torch

import torchacc_b200 as ta
from torchacc_b200.config import ConfigError
from torchacc_b200.parallel.mesh import Mesh, ProcessTopology


def test_defaults_and_predicates():
    c = ta.Config()
    c.validate()
    assert c.backend == "eager" and c.is_eager_backend() and not c.is_lazy_backend()
    assert c.dist.dp.size == 1 and not c.is_distributed_parallel() and not c.is_tracing_enabled()
    assert c.compute.dtype == torch.float32
    c.compute.bf16 = True
    assert c.compute.dtype == torch.bfloat16


def test_validation_errors():
    c = ta.Config()
    c.compute.fp16 = c.compute.bf16 = True
    with pytest.raises(ConfigError):
        c.validate()
    c = ta.Config()
    c.dist.pp.size = 2
    with pytest.raises(ConfigError):       # needs exactly one split point
        c.validate()
    c = ta.Config()
    c.dist.tp.size = 0
    with pytest.raises(ConfigError):
        c.validate()
    c = ta.Config()
    with pytest.raises(AttributeError):
        c.compute.no_such_field = 1
    c = ta.Config()
    c.dist.topology = ["dp", "dp"]
    with pytest.raises(ConfigError):
        c.validate()
    c = ta.Config()
    c.dataloader.buckets = [128, 64]
    with pytest.raises(ConfigError):
        c.validate()


def test_dp_size_inferred_from_world(monkeypatch):
    monkeypatch.setenv("WORLD_SIZE", "8")
    c = ta.Config()
    c.dist.fsdp.size = 2
    c.dist.tp.size = 2
    c.validate()
    assert c.dist.dp.size == 2 and c.is_distributed_parallel()
    c2 = ta.Config()
    c2.dist.fsdp.size = 3
    with pytest.raises(ConfigError):
        c2.validate()


def test_roundtrip_dict_json_yaml(tmp_path):
    c = ta.Config()
    c.compute.bf16 = True
    c.memory.gc = True
    c.memory.gc_cls = {"LlamaDecoderLayer"}
    c.dist.fsdp.wrap_layer_cls = {"A", "B"}
    c.dataloader.buckets = [128, 256]
    d = c.to_dict()
    c2 = ta.Config.from_dict(d)
    assert c2 == c
    p = tmp_path / "c.json"
    c.to_json(str(p))
    assert ta.Config.from_json(str(p)) == c
    import yaml
    assert ta.Config.from_yaml(yaml.safe_dump(d)) == c


def test_topology_groups():
    t = ProcessTopology(["dp", "fsdp", "tp"], [2, 2, 2])
    assert t.world_size() == 8
    assert t.get_rank(dp=1, fsdp=0, tp=1) == 5
    assert t.get_coord(6) == {"dp": 1, "fsdp": 1, "tp": 0}
    assert t.get_axis_comm_lists("tp") == [[0, 1], [2, 3], [4, 5], [6, 7]]
    assert t.get_axis_comm_lists("dp") == [[0, 4], [1, 5], [2, 6], [3, 7]]
    assert t.get_multi_axis_comm_lists(("dp", "fsdp")) == [[0, 2, 4, 6], [1, 3, 5, 7]]
    assert t.filter_match(dp=0, tp=1) == [1, 3]
    assert t.get_axis_comm_lists("pp") == [[r] for r in range(8)]


def test_mesh_accessors_without_process_group():
    m = Mesh(dp_num=2, pp_num=2, tp_num=2, topology=["dp", "pp", "tp"], rank=5, world_size=8, create_groups=False)
    assert (m.get_dp_rank(), m.get_pp_rank(), m.get_tp_rank()) == (1, 0, 1)
    assert m.get_stage_id() == 0 and m.is_first_stage() and not m.is_last_stage()
    assert m.stage_to_global(1) == 7
    assert m.get_tp_rank_groups() == [[0, 1], [2, 3], [4, 5], [6, 7]]
    assert m.get_dp_num() == 2 and m.get_fsdp_num() == 1 and m.get_sp_num() == 1
    with pytest.raises(ValueError):
        Mesh(dp_num=3, rank=0, world_size=8, create_groups=False)


def test_mesh_sp_axis_and_2d_split():
    m = Mesh(sp_num=4, dp_num=2, sp_mode="2d", ulysses_num=2, rank=6, world_size=8, create_groups=False)
    assert m.get_sp_num() == 4 and m.ulysses_num == 2 and m.ring_num == 2
    assert m.get_rank_groups("sp") == [[0, 1, 2, 3], [4, 5, 6, 7]]
    assert [4, 5] in m.get_rank_groups("ulysses") and [6, 7] in m.get_rank_groups("ulysses")
    assert [4, 6] in m.get_rank_groups("ring") and [5, 7] in m.get_rank_groups("ring")
    # 'sp' may appear in the topology (the reference raises here, dist/mesh.py:255-258)
    Mesh(sp_num=2, tp_num=2, topology=["sp", "tp"], rank=0, world_size=4, create_groups=False)


def test_api_surface():
    """Appendix A of SURVEY.md: the reference's public names exist."""
    for name in ["accelerate", "Config", "AsyncLoader", "amp", "sync", "lazy_device", "is_lazy_device", "is_lazy_tensor",
                 "fetch_gradients", "mark_dynamic", "save", "mark_step", "get_global_context", "__version__",
                 "accelerate_hf_trainer", "patch_qwen_model", "dist", "ops", "utils"]:
        assert hasattr(ta, name), name
    for name in ["world_size", "rank", "local_rank", "init_process_group", "init_nccl_context", "rendezvous", "Mesh",
                 "ParallelModule", "DataParallel", "FullyShardedDataParallel", "SpmdFullyShardedDataParallel",
                 "PipelineParallel", "DistributedParallel", "BACKEND_NAME", "EAGER_BACKEND_NAME", "fsdp", "pp", "tp"]:
        assert hasattr(ta.dist, name), name
    for name in ["flash_attn_xla", "flash_attn_varlen_xla", "flash_attn_varlen_qkvpacked_xla",
                 "spmd_flash_attn_varlen_xla", "flash_attn_varlen_position_ids_xla", "apply_liger_kernel",
                 "apply_liger_kernel_to_llama", "apply_liger_kernel_to_qwen2", "scaled_dot_product_attention"]:
        assert hasattr(ta.ops, name), name
    for name in ["ulysses", "ring_attention", "context_parallel_2d", "initialize_context_parallel",
                 "get_context_parallel_group", "get_inter_cp_process_group", "get_intra_cp_process_group",
                 "split_forward_gather_backward", "gather_forward_split_backward"]:
        assert hasattr(ta.ops.context_parallel, name), name
    assert hasattr(ta.utils.checkpoint, "gradient_checkpoint") and hasattr(ta.utils.checkpoint, "checkpoint_module")
    assert hasattr(ta.utils.cpu_offload, "get_cpu_offload_context")
    for name in ["patch_amp", "patch_fa", "patch_llama", "patch_qwen", "patch_autocast"]:
        assert hasattr(ta.utils.patch, name), name
    from torchacc_b200.parallel import state_dict_utils as U
    for name in ["consolidate_and_reshard_fsdp_model_dict", "consolidate_and_reshard_fsdp_optim_dict",
                 "consolidate_and_reshard_fsdp_checkpoint", "load_checkpoints", "save_checkpoints"]:
        assert hasattr(U, name), name
    from torchacc_b200.parallel.pp import schedule
    for name in ["PipeDreamFlushTrain", "PipeDreamFlushInfer", "LoadMicroBatch", "ForwardPass", "BackwardPass",
                 "SendActivation", "RecvActivation", "SendGrad", "RecvGrad", "ReduceGrads", "OptimizerStep"]:
        assert hasattr(schedule, name), name
