"""Optional-dependency probes (reference torchacc/utils/import_utils.py:6-11)."""
import functools
import importlib.util


@functools.lru_cache(maxsize=None)
def _has(name: str) -> bool:
    try:
        return importlib.util.find_spec(name) is not None
    except (ImportError, ValueError):
        return False


def is_torch_xla_available() -> bool:
    """Kept for API compatibility.  This framework never uses torch_xla, so the answer does not change behaviour."""
    return _has("torch_xla")


def is_transformers_available() -> bool:
    return _has("transformers")


def is_flash_attn_available() -> bool:
    return _has("flash_attn")
