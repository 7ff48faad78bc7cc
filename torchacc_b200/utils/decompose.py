"""``replace_decompose`` (reference torchacc/utils/decompose.py:23-128) rewrites torch's global decomposition tables
so in-place ops survive lazy-tensor tracing.  Nothing is traced here, so the function is an explicit no-op kept for
scripts that call it."""


def replace_decompose() -> None:
    return None
