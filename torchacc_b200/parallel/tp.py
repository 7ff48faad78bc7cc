"""Tensor parallelism -- placeholder header, implementation follows in this file (see below)."""
