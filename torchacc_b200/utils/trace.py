"""Symbolic tracing helpers for pipeline-stage construction (reference torchacc/utils/trace.py:21-175).

Pipeline parallelism in this framework does NOT require an fx trace: stages are cut on the module tree
(parallel/pp/partition.py).  ``trace`` is provided for models a user wants as a ``GraphModule``:

* HuggingFace models are traced at BLOCK level: decoder layers (``_no_split_modules``) and rotary-embedding modules
  stay leaf calls, the model is entered through a wrapper with explicit positional inputs (HF forwards are wrapped in
  ``**kwargs`` decorators that ``torch.fx`` cannot patch), and mask construction is skipped by selecting an attention
  implementation that builds no dense mask (our flash kernels mask in-kernel).  ``transformers.utils.fx`` -- what the
  reference calls (trace.py:21-77) -- no longer exists in transformers 5.
* everything else goes through ``torch.fx`` with the non-input arguments frozen to their defaults.

A failed trace raises; nothing falls back silently.

``lift_single_use_params`` is the pass the reference runs after ``split_module`` (trace.py:95-175): a parameter or
buffer fetched at the top level and consumed by exactly one stage moves INTO that stage, so every stage owns its
tensors (which is what the sharding engine and per-stage checkpoints need)."""
from __future__ import annotations

import inspect
import operator
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import torch
import torch.fx as fx
import torch.nn as nn


def is_getitem(node: fx.Node) -> bool:
    return node.op == "call_function" and node.target is operator.getitem


def is_output(node: fx.Node) -> bool:
    return node.op == "output"


def is_call_module(node: fx.Node) -> bool:
    return node.op == "call_module"


def get_concrete_args(model: nn.Module, input_names: List[str]) -> dict:
    """Arguments of ``model.forward`` not listed in ``input_names`` are frozen to their defaults."""
    sig = inspect.signature(model.forward)
    return {p.name: p.default for p in sig.parameters.values()
            if p.name not in input_names and p.default is not inspect.Parameter.empty
            and p.kind not in (inspect.Parameter.VAR_KEYWORD, inspect.Parameter.VAR_POSITIONAL)}


class BlockTracer(fx.Tracer):
    """``torch.fx`` tracer that keeps the given module classes (and torch.nn leaves) as single call_module nodes."""

    def __init__(self, leaf_classes: Sequence[type] = (), leaf_name_suffixes: Sequence[str] = ()):
        super().__init__()
        self.leaf_classes = tuple(leaf_classes)
        self.leaf_name_suffixes = tuple(leaf_name_suffixes)

    def is_leaf_module(self, m: nn.Module, qualname: str) -> bool:
        if self.leaf_classes and isinstance(m, self.leaf_classes):
            return True
        if self.leaf_name_suffixes and type(m).__name__.endswith(self.leaf_name_suffixes):
            return True
        return super().is_leaf_module(m, qualname)


def _make_entry(model: nn.Module, input_names: Sequence[str], output_attr: Optional[str]) -> nn.Module:
    """A module whose forward takes exactly ``input_names`` positionally and calls ``model`` by keyword."""
    args = ", ".join(input_names)
    kwargs = ", ".join(f"{n}={n}" for n in input_names)
    src = (f"def forward(self, {args}):\n"
           f"    out = self.model({kwargs})\n"
           f"    return out" + (f".{output_attr}" if output_attr else "") + "\n")
    ns: dict = {}
    exec(src, ns)            # generated from identifiers validated below
    cls = type("TraceEntry", (nn.Module,), {"forward": ns["forward"]})
    entry = cls()
    entry.model = model
    return entry


def _hf_leaf_classes(model: nn.Module) -> Tuple[type, ...]:
    names = set(getattr(model, "_no_split_modules", None) or ())
    return tuple({type(m) for m in model.modules() if type(m).__name__ in names})


def is_hf_model(model: nn.Module) -> bool:
    try:
        from transformers import PreTrainedModel
    except ImportError:
        return False
    return isinstance(model, PreTrainedModel)


def hf_trace_entry(model: nn.Module, input_names: Sequence[str], output_attr: Optional[str] = "logits"):
    """Prepare a HuggingFace model for block-level tracing: returns ``(entry module, leaf classes, leaf name suffixes)``.
    The entry takes ``input_names`` positionally and addresses the original modules under the ``model.`` prefix.  Side
    effects on ``model.config`` (kept: the traced graph relies on them): attention implementation ``"torchacc_b200"`` (no
    dense mask is built; our kernels mask in-kernel) and ``use_cache = False``."""
    for n in input_names:
        if not n.isidentifier():
            raise ValueError(f"input name {n!r} is not an identifier")
    from .patch import patch_fa
    patch_fa()                                            # registers the "torchacc_b200" attention interface
    model.config._attn_implementation = "torchacc_b200"
    if getattr(model.config, "use_cache", None) is not None:
        model.config.use_cache = False
    # norm modules stay leaves too: their forward may be patched onto our fused kernels (ops/liger.py), which are opaque to fx
    return _make_entry(model, list(input_names), output_attr), _hf_leaf_classes(model), ("RotaryEmbedding", "RMSNorm", "LayerNorm")


def trace(model: nn.Module, input_names: Optional[List[str]] = None, leaf_classes: Iterable[type] = (),
          output_attr: Optional[str] = None) -> fx.GraphModule:
    """GraphModule of ``model`` with ``input_names`` as placeholders.  HF models: block-level (see module docstring);
    ``output_attr`` selects a field of the HF output object (default ``"logits"``).  The traced HF graph addresses
    the original modules under the ``model.`` prefix."""
    sig_names = [n for n, p in inspect.signature(model.forward).parameters.items()
                 if p.kind not in (inspect.Parameter.VAR_KEYWORD, inspect.Parameter.VAR_POSITIONAL)]
    input_names = list(input_names) if input_names else sig_names[:1]
    if not is_hf_model(model):
        tracer = BlockTracer(tuple(leaf_classes))
        graph = tracer.trace(model, concrete_args=get_concrete_args(model, input_names))
        return fx.GraphModule(model, graph)

    saved_impl = model.config._attn_implementation
    entry, hf_leaves, suffixes = hf_trace_entry(model, input_names, output_attr if output_attr is not None else "logits")
    tracer = BlockTracer(tuple(leaf_classes) + hf_leaves, suffixes)
    try:
        graph = tracer.trace(entry)
    except Exception as e:
        model.config._attn_implementation = saved_impl
        raise RuntimeError(
            f"block-level fx trace of {type(model).__name__} failed ({type(e).__name__}: {e}); pipeline stages do not "
            f"need a trace -- cut the module tree with dist.pp.split_points instead") from e
    return fx.GraphModule(entry, graph)


def _placeholder_for(callee: fx.GraphModule, user: fx.Node, node: fx.Node) -> List[fx.Node]:
    """Placeholders of ``callee`` fed by ``node`` in the call ``user`` (positional or keyword)."""
    phs = [n for n in callee.graph.nodes if n.op == "placeholder"]
    found = [phs[i] for i, a in enumerate(user.args) if a is node]
    by_name = {p.target: p for p in phs}
    found += [by_name[k] for k, a in user.kwargs.items() if a is node and k in by_name]
    return found


def lift_single_use_params(split: fx.GraphModule, qualname_map: Optional[Dict[str, str]] = None) -> Dict[str, str]:
    """Move every top-level ``get_attr`` tensor whose only consumer is one ``call_module`` into that submodule.

    ``split`` is the result of ``torch.fx.passes.split_module.split_module``.  Returns (and updates in place, if
    given) the map ``new qualified name -> original qualified name`` so checkpoints keep their original keys."""
    qualname_map = {} if qualname_map is None else qualname_map
    moved: List[Tuple[nn.Module, str]] = []
    touched = set()
    for node in list(split.graph.nodes):
        if node.op != "get_attr" or len(node.users) != 1:
            continue
        user = next(iter(node.users))
        if user.op != "call_module":
            continue
        callee = split.get_submodule(user.target)
        if not isinstance(callee, fx.GraphModule):
            continue
        owner_path, _, leaf = node.target.rpartition(".")
        owner = split.get_submodule(owner_path) if owner_path else split
        value = getattr(owner, leaf)
        if not isinstance(value, torch.Tensor):
            continue
        phs = _placeholder_for(callee, user, node)
        if not phs:
            continue
        new_name = "lifted_" + node.target.replace(".", "_")
        if hasattr(callee, new_name):
            raise RuntimeError(f"{user.target} already has an attribute {new_name}")
        if leaf in owner._buffers:
            callee.register_buffer(new_name, value, persistent=leaf not in owner._non_persistent_buffers_set)
        elif isinstance(value, nn.Parameter):
            callee.register_parameter(new_name, value)
        else:
            setattr(callee, new_name, value)
        for ph in phs:
            with callee.graph.inserting_before(ph):
                fetched = callee.graph.get_attr(new_name)
            ph.replace_all_uses_with(fetched)
        # drop the argument from the call; the callee's signature shrinks accordingly
        drop = set(id(p) for p in phs)
        all_phs = [n for n in callee.graph.nodes if n.op == "placeholder"]
        keep_pos = [i for i, p in enumerate(all_phs) if id(p) not in drop]
        user.args = tuple(a for i, a in enumerate(user.args) if i in keep_pos and a is not node)
        user.kwargs = {k: a for k, a in user.kwargs.items() if a is not node}
        for ph in phs:
            callee.graph.erase_node(ph)
        split.graph.erase_node(node)
        key = f"{user.target}.{new_name}"
        qualname_map[key] = qualname_map.pop(node.target, node.target)
        moved.append((owner, leaf))
        touched.add(user.target)
    for name in touched:
        callee = split.get_submodule(name)
        callee.graph.lint()
        callee.recompile()
    still_used = {n.target for n in split.graph.nodes if n.op == "get_attr"}
    for owner, leaf in moved:
        full = [k for k, m in split.named_modules() if m is owner]
        path = (full[0] + "." if full and full[0] else "") + leaf
        if path not in still_used and hasattr(owner, leaf):
            delattr(owner, leaf)
    split.delete_all_unused_submodules()
    split.graph.lint()
    split.recompile()
    return qualname_map
