"""HuggingFace integration: kernel patches (reference torchacc/ops/liger.py, utils/patch.py) keep HF Llama / Qwen2
numerics, load_hf_state_dict maps an HF checkpoint onto the native model, accelerate() accepts an HF model."""
import copy

import pytest
import torch

transformers = pytest.importorskip("transformers")


def _hf_llama(attn="eager"):
    from transformers import LlamaConfig, LlamaForCausalLM
    cfg = LlamaConfig(vocab_size=160, hidden_size=64, intermediate_size=128, num_hidden_layers=2,
                      num_attention_heads=4, num_key_value_heads=2, max_position_embeddings=64,
                      attn_implementation=attn)
    torch.manual_seed(0)
    return LlamaForCausalLM(cfg)


def _hf_qwen2():
    from transformers import Qwen2Config, Qwen2ForCausalLM
    cfg = Qwen2Config(vocab_size=160, hidden_size=64, intermediate_size=128, num_hidden_layers=2,
                      num_attention_heads=4, num_key_value_heads=2, max_position_embeddings=64,
                      attn_implementation="eager")
    torch.manual_seed(0)
    return Qwen2ForCausalLM(cfg)


@pytest.fixture
def restore_hf_classes():
    from transformers.models.llama import modeling_llama as ml
    from transformers.models.qwen2 import modeling_qwen2 as mq
    saved = [(c, c.forward) for c in (ml.LlamaRMSNorm, ml.LlamaMLP, ml.LlamaForCausalLM, mq.Qwen2RMSNorm, mq.Qwen2MLP,
                                      mq.Qwen2ForCausalLM)]
    yield
    for c, f in saved:
        c.forward = f
        if hasattr(c, "_tb_lce_patched"):
            del c._tb_lce_patched
    from torchacc_b200.utils.patch import unpatch_all
    unpatch_all()


@pytest.mark.parametrize("family", ["llama", "qwen2"])
def test_liger_style_patches_preserve_loss_and_grads(family, restore_hf_classes):
    import torchacc_b200 as ta
    model = _hf_llama() if family == "llama" else _hf_qwen2()
    ids = torch.randint(0, 160, (2, 24), generator=torch.Generator().manual_seed(3))
    ref = model(input_ids=ids, labels=ids)
    ref.loss.backward()
    ref_grads = {n: p.grad.clone() for n, p in model.named_parameters()}
    model.zero_grad()
    (ta.ops.apply_liger_kernel_to_llama if family == "llama" else ta.ops.apply_liger_kernel_to_qwen2)()
    out = model(input_ids=ids, labels=ids)
    assert out.logits is None                     # fused linear + cross-entropy: no [T, V] logits
    assert abs(float(out.loss) - float(ref.loss)) < 1e-4
    out.loss.backward()
    for n, p in model.named_parameters():
        assert torch.allclose(p.grad, ref_grads[n], atol=2e-5, rtol=1e-3), n
    # without labels the original forward (with logits) is used
    assert model(input_ids=ids).logits.shape == (2, 24, 160)


def test_patch_fa_registers_attention_interface(restore_hf_classes):
    """The 'flash_attention_2' entry of HF's attention registry is served by our attention op."""
    from torchacc_b200.utils.patch import patch_fa
    assert patch_fa()
    from transformers.modeling_utils import ALL_ATTENTION_FUNCTIONS
    fn = ALL_ATTENTION_FUNCTIONS["torchacc_b200"]
    g = torch.Generator().manual_seed(0)
    q = torch.randn(2, 4, 16, 16, generator=g)        # HF layout [B, H, S, D]
    k = torch.randn(2, 2, 16, 16, generator=g)
    v = torch.randn(2, 2, 16, 16, generator=g)

    class M:
        is_causal = True
    out, _ = fn(M(), q, k, v, None, scaling=16 ** -0.5)
    kk, vv = k.repeat_interleave(2, 1), v.repeat_interleave(2, 1)
    ref = torch.nn.functional.scaled_dot_product_attention(q, kk, vv, is_causal=True).transpose(1, 2)
    assert torch.allclose(out, ref, atol=1e-5)


def test_native_model_loads_hf_checkpoint():
    from torchacc_b200.models import build_llama
    hf = _hf_llama()
    native = build_llama("tiny", hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=4,
                         num_key_value_heads=2, head_dim=16, vocab_size=160, max_position_embeddings=64,
                         rope_theta=hf.config.rope_parameters["rope_theta"] if hasattr(hf.config, "rope_parameters")
                         else hf.config.rope_theta, rms_norm_eps=hf.config.rms_norm_eps)
    native.load_hf_state_dict(hf.state_dict())
    ids = torch.randint(0, 160, (2, 24), generator=torch.Generator().manual_seed(3))
    a = hf(input_ids=ids, labels=ids)
    b = native(ids, labels=ids, return_logits=True)
    assert abs(float(a.loss) - float(b["loss"])) < 1e-4
    assert torch.allclose(a.logits, b["logits"], atol=1e-4)
    # and back
    sd = native.to_hf_state_dict()
    for k, v in hf.state_dict().items():
        assert torch.equal(sd[k], v), k


def test_accelerate_hf_model_single_process(restore_hf_classes):
    import torchacc_b200 as ta
    model = _hf_llama()
    ref = copy.deepcopy(model)
    cfg = ta.Config()
    cfg.memory.gc = True
    cfg.memory.gc_cls = {"LlamaDecoderLayer"}
    m = ta.accelerate(model, config=cfg)
    ids = torch.randint(0, 160, (2, 24), generator=torch.Generator().manual_seed(3))
    out = m(input_ids=ids, labels=ids)
    loss = out.loss if hasattr(out, "loss") else out["loss"]
    r = ref(input_ids=ids, labels=ids)
    assert abs(float(loss) - float(r.loss)) < 1e-4
    loss.backward()


def _hf_fsdp_worker(rank, world):
    import torchacc_b200 as ta
    model = _hf_llama()
    ref = copy.deepcopy(model)
    cfg = ta.Config()
    cfg.dist.fsdp.size = world
    cfg.dist.fsdp.wrap_layer_cls = {"LlamaDecoderLayer"}
    cfg.memory.gc = True
    m = ta.accelerate(model, config=cfg)
    opt = torch.optim.SGD(m.parameters(), lr=0.1)
    ref_opt = torch.optim.SGD(ref.parameters(), lr=0.1)
    ids = torch.randint(0, 160, (4, 24), generator=torch.Generator().manual_seed(3))
    local = ids.chunk(world)[rank]
    for _ in range(2):
        out = m(input_ids=local, labels=local)
        (out.loss if hasattr(out, "loss") else out["loss"]).backward()
        opt.step()
        m.zero_grad()
        ref(input_ids=ids, labels=ids).loss.backward()
        ref_opt.step()
        ref_opt.zero_grad()
    full = m._inner_engine_module().full_state_dict(rank0_only=False)
    for n, p in ref.named_parameters():
        assert torch.allclose(full[n], p.detach(), atol=1e-4, rtol=1e-3), (n, float((full[n] - p).abs().max()))


def test_accelerate_hf_model_fsdp_two_ranks():
    """An unmodified HuggingFace model through accelerate(): class patches + FSDP engine + GC == single process."""
    from dist_utils import run_distributed
    run_distributed(_hf_fsdp_worker, 2)


def test_native_gpt2_matches_hf_gpt2():
    """Second native family: same weights -> same loss / logits as HuggingFace GPT-2 (vocabulary 131 is padded to 256 in
    our embedding; the padding rows must not take part in the softmax)."""
    from transformers import GPT2Config as HFConfig, GPT2LMHeadModel as HFGPT2
    from torchacc_b200.models import build_gpt2
    torch.manual_seed(0)
    hf = HFGPT2(HFConfig(vocab_size=131, n_positions=64, n_embd=64, n_layer=2, n_head=4, attn_pdrop=0.0, embd_pdrop=0.0,
                         resid_pdrop=0.0)).eval()
    ours = build_gpt2("gpt2-tiny", vocab_size=131, n_positions=64, n_embd=64, n_layer=2, n_head=4)
    assert ours.config.padded_vocab == 256
    ours.load_hf_state_dict(hf.state_dict())
    with torch.no_grad():
        ours.wte.weight[131:].normal_(0, 1.0)         # garbage in the padding rows must not matter
    ids = torch.randint(0, 131, (2, 24), generator=torch.Generator().manual_seed(4))
    a = hf(input_ids=ids, labels=ids)
    b = ours(ids, labels=ids, return_logits=True)
    assert torch.allclose(a.logits, b["logits"], atol=2e-4), float((a.logits - b["logits"]).abs().max())
    assert abs(float(a.loss) - float(b["loss"])) < 1e-4, (float(a.loss), float(b["loss"]))


def test_accelerate_hf_trainer_patches_drive_ten_steps(restore_hf_classes):
    """reference core/accelerate_hf_trainer.py:52-77.  `transformers.Trainer` cannot be constructed offline (it needs
    the `accelerate` package), so the three patched Trainer methods are driven in the order Trainer's inner loop calls
    them (trainer.py: _wrap_model -> create_optimizer -> per step: forward/backward, _clip_grad_norm, optimizer.step,
    optimizer.zero_grad) on a stand-in that carries the attributes those methods read."""
    import types
    import torchacc_b200 as ta
    from transformers import Trainer
    assert ta.accelerate_hf_trainer(True)
    try:
        args = types.SimpleNamespace(bf16=False, fp16=False, gradient_checkpointing=True, fsdp="full_shard",
                                     fsdp_config={"transformer_layer_cls_to_wrap": ["LlamaDecoderLayer"]},
                                     learning_rate=1e-2, adam_beta1=0.9, adam_beta2=0.999, adam_epsilon=1e-8,
                                     weight_decay=0.0, max_grad_norm=1.0)
        me = types.SimpleNamespace(args=args, optimizer=None, model=_hf_llama(), model_wrapped=None)
        model = Trainer._wrap_model(me, me.model)
        assert getattr(model, "_tb_accelerated", False) and me.model_wrapped is model
        assert Trainer._wrap_model(me, model) is model            # idempotent, like HF's own wrapper
        opt = Trainer.create_optimizer(me)
        assert type(opt).__name__ == "FusedAdamW" and Trainer.create_optimizer(me) is opt
        ids = torch.randint(0, 160, (2, 24), generator=torch.Generator().manual_seed(3))
        losses, norms = [], []
        for _ in range(10):
            loss = model(input_ids=ids, labels=ids).loss
            loss.backward()
            norms.append(float(Trainer._clip_grad_norm(me, model)))   # routed through the sharding engine
            opt.step()
            opt.zero_grad()
            losses.append(float(loss))
        assert losses[-1] < losses[0] - 0.5, losses
        assert all(n > 0 for n in norms)
        assert model.engine.last_clip_coef is not None             # the engine computed the coefficient
    finally:
        ta.accelerate_hf_trainer(False)
    assert Trainer._clip_grad_norm.__qualname__.startswith("Trainer.")


def test_accelerate_accepts_an_unpatched_hf_family_gpt2(restore_hf_classes):
    """Families without class-level kernel patches (here HF GPT-2, the reference benchmark's default model,
    benchmarks/transformer.py:37) run through accelerate() as they are: same loss, gradients on the flat shards."""
    import torchacc_b200 as ta
    from transformers import GPT2Config, GPT2LMHeadModel
    gc = GPT2Config(vocab_size=160, n_positions=64, n_embd=64, n_layer=2, n_head=4, resid_pdrop=0.0, embd_pdrop=0.0,
                    attn_pdrop=0.0)
    torch.manual_seed(0)
    ref = GPT2LMHeadModel(gc)
    model = copy.deepcopy(ref)
    ids = torch.randint(0, 160, (2, 16), generator=torch.Generator().manual_seed(2))
    want = ref(input_ids=ids, labels=ids).loss
    cfg = ta.Config()
    cfg.compute.bf16 = False
    cfg.dist.fsdp.wrap_layer_cls = {"GPT2Block"}
    model = ta.accelerate(model, config=cfg)
    out = model(input_ids=ids, labels=ids)
    assert abs(float(out.loss) - float(want)) < 1e-4
    out.loss.backward()
    assert all(g is not None and torch.isfinite(g).all() for g in model.engine.grads())
