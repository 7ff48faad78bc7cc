"""Stand-in for the HuggingFace `accelerate` package, which is not installable offline in this image.

The UNMODIFIED reference (baseline/_ref/torchacc) does `import accelerate` at import time
(torchacc/core/accelerate_hf_trainer.py:4) but only touches its attributes inside `accelerate_hf_trainer(True)`,
which the benchmark's reference arm never calls.  This empty module satisfies the import; it is a dependency
stub, not a change to the reference."""
__version__ = "0.0.0"
