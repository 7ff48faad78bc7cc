"""LLM helpers (reference torchacc/llm/)."""
from .qwen_patch import patch_qwen_model, rewrite_forward

__all__ = ["patch_qwen_model", "rewrite_forward"]
