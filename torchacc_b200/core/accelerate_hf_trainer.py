"""HuggingFace ``Trainer`` integration (reference torchacc/core/accelerate_hf_trainer.py:21-77).

The reference makes ``transformers``/``accelerate`` believe a TPU is present so the Trainer takes its XLA code
path.  That trick has no meaning without XLA.  Here ``accelerate_hf_trainer(True)`` patches
``transformers.Trainer`` so that (a) the model is wrapped by ``torchacc_b200.accelerate`` with a Config derived
from the TrainingArguments (bf16/fp16, fsdp settings, gradient checkpointing), (b) the optimizer is built over
the flat sharded parameters (FusedAdamW when the requested optimizer is AdamW) and (c) gradient clipping goes
through the engine.  Everything else (logging, LR schedule, checkpoint cadence) stays HF's.
"""
from __future__ import annotations

import os

from ..utils.logger import logger

_ENABLED = False
_ORIG = {}


def _config_from_args(args):
    from ..config import Config
    cfg = Config()
    cfg.compute.bf16 = bool(getattr(args, "bf16", False))
    cfg.compute.fp16 = bool(getattr(args, "fp16", False))
    cfg.memory.gc = bool(getattr(args, "gradient_checkpointing", False))
    fsdp_cfg = getattr(args, "fsdp_config", None) or {}
    xla = fsdp_cfg.get("xla_fsdp_settings", {}) if isinstance(fsdp_cfg, dict) else {}
    wrap = fsdp_cfg.get("transformer_layer_cls_to_wrap") or fsdp_cfg.get("fsdp_transformer_layer_cls_to_wrap")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if getattr(args, "fsdp", None):
        cfg.dist.fsdp.size = world
        if wrap:
            cfg.dist.fsdp.wrap_layer_cls = set(wrap if isinstance(wrap, (list, tuple, set)) else [wrap])
        cfg.dist.fsdp.flatten_parameters = bool(xla.get("flatten_parameters", True))
    return cfg


def accelerate_hf_trainer(enable: bool = True) -> bool:
    """Install (or remove) the Trainer patches.  Returns False when ``transformers`` is unavailable."""
    global _ENABLED
    try:
        import transformers  # noqa: F401
        from transformers import Trainer
    except Exception as e:  # pragma: no cover
        logger.warning("accelerate_hf_trainer: transformers is not importable (%s)", e)
        return False
    if enable and not _ENABLED:
        _ORIG["_wrap_model"] = Trainer._wrap_model
        _ORIG["create_optimizer"] = Trainer.create_optimizer

        def _wrap_model(self, model, training=True, dataloader=None):
            if getattr(model, "_tb_accelerated", False) or not training:
                return model
            from ..accelerate import accelerate
            wrapped = accelerate(model, config=_config_from_args(self.args))
            object.__setattr__(wrapped, "_tb_accelerated", True)
            self.model_wrapped = wrapped
            return wrapped

        def create_optimizer(self):
            if self.optimizer is None:
                from ..ops.optim import FusedAdamW
                model = getattr(self, "model_wrapped", None) or self.model
                if not getattr(model, "_tb_accelerated", False):
                    model = _wrap_model(self, model)
                a = self.args
                self.optimizer = FusedAdamW(model.parameters(), lr=a.learning_rate, betas=(a.adam_beta1, a.adam_beta2),
                                            eps=a.adam_epsilon, weight_decay=a.weight_decay)
            return self.optimizer

        def _clip_grad_norm(self, model):
            # transformers >= 4.5x calls this once per optimizer step (trainer.py: `grad_norm = self._clip_grad_norm(model)`)
            m = model if getattr(model, "_tb_accelerated", False) else getattr(self, "model_wrapped", model)
            if hasattr(m, "clip_grad_norm_") and getattr(m, "_tb_accelerated", False):
                return m.clip_grad_norm_(self.args.max_grad_norm)     # sharded norm + device-side coefficient
            return _ORIG["_clip_grad_norm"](self, model)

        Trainer._wrap_model = _wrap_model
        Trainer.create_optimizer = create_optimizer
        if hasattr(Trainer, "_clip_grad_norm"):
            _ORIG["_clip_grad_norm"] = Trainer._clip_grad_norm
            Trainer._clip_grad_norm = _clip_grad_norm
        _ENABLED = True
    elif not enable and _ENABLED:
        Trainer._wrap_model = _ORIG.pop("_wrap_model")
        Trainer.create_optimizer = _ORIG.pop("create_optimizer")
        if "_clip_grad_norm" in _ORIG:
            Trainer._clip_grad_norm = _ORIG.pop("_clip_grad_norm")
        _ENABLED = False
    return True
