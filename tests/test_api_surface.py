"""Public API checklist: every name a TorchAcc user reaches through `import torchacc as ta`
exists in `import torchacc_b200 as ta` with a compatible call shape."""
import inspect

import pytest
import torch

import torchacc_b200 as ta


def _has(obj, dotted):
    for part in dotted.split("."):
        assert hasattr(obj, part), f"missing {dotted} (at {part})"
        obj = getattr(obj, part)
    return obj


TOP = ["accelerate", "Config", "AsyncLoader", "amp.GradScaler", "sync", "lazy_device", "is_lazy_device", "is_lazy_tensor",
       "fetch_gradients", "mark_dynamic", "save", "mark_step", "get_global_context", "__version__",
       "accelerate_hf_trainer", "patch_qwen_model"]
DIST = ["world_size", "rank", "local_rank", "init_process_group", "init_nccl_context", "rendezvous", "Mesh",
        "ParallelModule", "DataParallel", "FullyShardedDataParallel", "SpmdFullyShardedDataParallel", "PipelineParallel",
        "DistributedParallel", "BACKEND_NAME", "EAGER_BACKEND_NAME", "fsdp", "pp", "tp"]
OPS = ["flash_attn_xla", "flash_attn_varlen_xla", "flash_attn_varlen_qkvpacked_xla", "spmd_flash_attn_varlen_xla",
       "flash_attn_varlen_position_ids_xla", "apply_liger_kernel", "apply_liger_kernel_to_llama",
       "apply_liger_kernel_to_qwen2", "scaled_dot_product_attention"]
CP = ["ulysses", "ring_attention", "context_parallel_2d", "initialize_context_parallel", "get_context_parallel_group",
      "get_inter_cp_process_group", "get_intra_cp_process_group", "split_forward_gather_backward",
      "gather_forward_split_backward"]
UTILS = ["checkpoint.gradient_checkpoint", "checkpoint.checkpoint_module", "cpu_offload.get_cpu_offload_context", "logger",
         "patch.patch_amp", "patch.patch_fa", "patch.patch_llama", "patch.patch_qwen", "patch.patch_autocast",
         "trace.trace"]
SDU = ["consolidate_and_reshard_fsdp_model_dict", "consolidate_and_reshard_fsdp_optim_dict",
       "consolidate_and_reshard_fsdp_checkpoint", "load_checkpoints", "save_checkpoints"]
MESH = [f"get_{a}_{b}" for a in ("dp", "pp", "tp", "fsdp") for b in ("rank", "num", "proc_group", "rank_groups")] + \
       ["get_stage_id", "is_first_stage", "is_last_stage", "stage_to_global", "get_sp_num", "get_global_rank",
        "get_world_size"]


@pytest.mark.parametrize("name", TOP)
def test_top_level(name):
    _has(ta, name)


@pytest.mark.parametrize("name", DIST)
def test_dist(name):
    _has(ta.dist, name)


@pytest.mark.parametrize("name", OPS)
def test_ops(name):
    _has(ta.ops, name)


@pytest.mark.parametrize("name", CP)
def test_context_parallel(name):
    _has(ta.ops.context_parallel, name)


@pytest.mark.parametrize("name", UTILS)
def test_utils(name):
    _has(ta.utils, name)


@pytest.mark.parametrize("name", SDU)
def test_state_dict_utils(name):
    _has(ta.dist.state_dict_utils, name)


def test_mesh_getters_and_pp_modules():
    for n in MESH:
        assert hasattr(ta.dist.Mesh, n), n
    _has(ta.dist.pp, "PipelineParallel")
    _has(ta.dist.pp, "preprocess_config")
    import torchacc_b200.parallel.pp.schedule as sched
    for n in ("PipeSchedule", "PipeDreamFlushTrain", "PipeDreamFlushInfer", "PipeInstruction", "Algo", "ForwardPass", "BackwardPass", "SendActivation", "RecvActivation",
              "SendGrad", "RecvGrad", "OptimizerStep"):
        assert hasattr(sched, n), n
    from torchacc_b200.parallel.pp.executor import PipeExecutor
    assert hasattr(PipeExecutor, "reset_activation_shape")


def test_config_fields_and_methods():
    c = ta.Config()
    for path in ("backend", "compute.fp16", "compute.bf16", "compute.acc_scaled_dot_attn",
                 "compute.disable_kernel_patches", "compute.fp8", "memory.gc", "memory.gc_cls", "memory.gc_cnt",
                 "dataloader.buckets", "dataloader.max_length", "dataloader.num_buckets", "dataloader.pad_value_dict",
                 "dist.dp.size", "dist.tp.size", "dist.pp.size", "dist.pp.num_micro_batches", "dist.pp.input_names",
                 "dist.pp.split_points", "dist.pp.broadcast_loss", "dist.fsdp.size", "dist.fsdp.wrap_layer_cls",
                 "dist.fsdp.flatten_parameters", "dist.fsdp.sync_module_states", "dist.fsdp.use_spmd",
                 "dist.fsdp.shard_output_callable", "dist.sp.size", "dist.topology"):
        _has(c, path)
    for m in ("validate", "get_mesh", "is_distributed_parallel", "is_tracing_enabled", "is_lazy_backend",
              "is_eager_backend"):
        assert callable(getattr(c, m)), m


def test_call_shapes():
    sig = inspect.signature(ta.accelerate)
    assert list(sig.parameters)[:3] == ["model", "dataloader", "config"]
    sig = inspect.signature(ta.AsyncLoader.__init__)
    for p in ("loader", "device", "buckets", "max_length", "num_buckets", "pad_value_dict"):
        assert p in sig.parameters, p
    sig = inspect.signature(ta.amp.GradScaler.__init__)
    for p in ("init_scale", "growth_factor", "backoff_factor", "growth_interval", "enabled"):
        assert p in sig.parameters, p
    sig = inspect.signature(ta.ops.flash_attn_xla)
    for p in ("q", "k", "v", "dropout_p", "softmax_scale", "causal", "window_size", "alibi_slopes", "deterministic",
              "return_attn_probs"):
        assert p in sig.parameters, p
    for fn, params in ((ta.ops.context_parallel.ulysses, ("q", "k", "v")),
                       (ta.ops.context_parallel.ring_attention, ("q", "k", "v")),
                       (ta.ops.context_parallel.context_parallel_2d, ("q", "k", "v"))):
        got = list(inspect.signature(fn).parameters)
        assert tuple(got[:3]) == params, (fn.__name__, got)
    m = ta.accelerate(torch.nn.Linear(4, 4))
    for attr in ("device", "forward_backward", "clip_grad_norm_", "sharded_optim_state_dict", "full_optim_state_dict",
                 "optim_state_dict_to_load"):
        assert hasattr(m, attr), attr
