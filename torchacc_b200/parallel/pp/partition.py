"""Cutting a model into pipeline stages (reference torchacc/dist/pp/utils.py:12-350 + pipeline.py:70-92).

Two routes produce the same artefact -- a list of ``StageSpec`` (one per stage) saying which values a stage
receives from its predecessor, which it loads from the data batch, and which it must send on:

1. **native protocol**: a model that defines ``pipeline_stages(split_points) -> List[nn.Module]`` (our Llama / GPT-2)
   is cut on its module tree, no tracing involved;
2. **fx route** for arbitrary ``nn.Module``s: a tagging tracer records, for every graph node, the module path it was
   created under; nodes are assigned to stages by "the split happens *before* the named module runs"
   (reference config.py:175-178), ``torch.fx.passes.split_module`` builds the per-stage GraphModules and a
   liveness pass over the top-level graph decides what crosses each boundary.  A value produced in stage i and
   consumed in stage j > i+1 is simply *live* across boundaries i..j-1, so intermediate stages forward it
   unchanged -- the reference implements the same behaviour by rewriting every stage graph
   (``_propagate_output``, pp/utils.py:85-239).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any, Callable, Dict, List, Optional, Sequence, Union

import torch.fx as fx
import torch.nn as nn
from torch.fx.passes.split_module import split_module


@dataclass
class StageSpec:
    index: int
    module: nn.Module                       # called as module(**inputs) -> Dict[str, Any] (or tensor / tuple)
    recv_names: List[str] = field(default_factory=list)   # values arriving from the previous stage
    load_names: List[str] = field(default_factory=list)   # values taken from the (micro-)batch kwargs
    send_names: List[str] = field(default_factory=list)   # values shipped to the next stage
    output_names: List[str] = field(default_factory=list) # last stage only: names forming the model output
    rebuild_output: Optional[Callable[[Dict[str, Any]], Any]] = None


def names_of_split_points(model: nn.Module, split_points: Sequence[Union[str, nn.Module]]) -> List[str]:
    """Module objects -> qualified names (reference pipeline.py:13-25 ``preprocess_config``)."""
    by_id = {id(m): n for n, m in model.named_modules()}
    out = []
    for p in split_points:
        if isinstance(p, str):
            if p not in dict(model.named_modules()):
                raise ValueError(f"split point '{p}' is not a submodule of the model")
            out.append(p)
        else:
            if id(p) not in by_id:
                raise ValueError("split point module is not part of the model")
            out.append(by_id[id(p)])
    return out


# ------------------------------------------------------------------------------------------------------------
# fx route
# ------------------------------------------------------------------------------------------------------------
class _TaggingTracer(fx.Tracer):
    """Records the module path under which each node is created (``node.meta['mod_path']``)."""

    def __init__(self, leaf_paths: Sequence[str] = (), leaf_classes: Sequence[type] = (),
                 leaf_name_suffixes: Sequence[str] = ()):
        super().__init__()
        self._stack: List[str] = []
        self._leaf_paths = set(leaf_paths)
        self._leaf_classes = tuple(leaf_classes)
        self._leaf_suffixes = tuple(leaf_name_suffixes)

    def is_leaf_module(self, m, qualname):
        if qualname in self._leaf_paths or (self._leaf_classes and isinstance(m, self._leaf_classes)):
            return True
        if self._leaf_suffixes and type(m).__name__.endswith(self._leaf_suffixes):
            return True
        return super().is_leaf_module(m, qualname)

    def call_module(self, m, forward, args, kwargs):
        self._stack.append(self.path_of_module(m))
        try:
            return super().call_module(m, forward, args, kwargs)
        finally:
            self._stack.pop()

    def create_node(self, *a, **k):
        node = super().create_node(*a, **k)
        node.meta["mod_path"] = self._stack[-1] if self._stack else ""
        if node.op == "call_module" and not node.meta["mod_path"]:
            node.meta["mod_path"] = str(node.target)
        return node


def _under(path: str, root: str) -> bool:
    return path == root or path.startswith(root + ".")


def trace_and_split(model: nn.Module, split_names: List[str], input_names: Optional[List[str]]) -> List[StageSpec]:
    import inspect
    from ...utils.trace import hf_trace_entry, is_hf_model, lift_single_use_params
    if is_hf_model(model):
        # HuggingFace forwards cannot be traced whole (kwargs decorators, cache / mask control flow): block-level trace
        # through an explicit-input entry module (utils/trace.py).  The last stage returns the logits; the loss is the
        # user's ``output_fn(logits, labels=...)`` like in the reference's pipeline examples.
        input_names = list(input_names) if input_names else ["input_ids"]
        root, leaf_classes, leaf_suffixes = hf_trace_entry(model, input_names, "logits")
        tracer = _TaggingTracer(leaf_classes=leaf_classes, leaf_name_suffixes=leaf_suffixes)
        graph = tracer.trace(root)
        split_names = ["model." + n for n in split_names]
    else:
        sig = inspect.signature(model.forward)
        input_names = input_names or [n for n, p in sig.parameters.items() if p.default is inspect.Parameter.empty]
        concrete = {p.name: p.default for p in sig.parameters.values()
                    if p.name not in input_names and p.default is not inspect.Parameter.empty}
        root = model
        tracer = _TaggingTracer()
        graph = tracer.trace(model, concrete_args=concrete)
    gm = fx.GraphModule(root, graph)
    # stage assignment: bump the stage when the first node belonging to the next split module appears
    stage_of: Dict[fx.Node, int] = {}
    cur, nxt = 0, 0
    for node in gm.graph.nodes:
        path = node.meta.get("mod_path", "")
        if node.op == "call_module":
            path = str(node.target)
        while nxt < len(split_names) and path and _under(path, split_names[nxt]):
            cur, nxt = cur + 1, nxt + 1
        stage_of[node] = cur
    if nxt != len(split_names):
        raise ValueError(f"split points {split_names[nxt:]} were never reached while tracing the model")
    num_stages = len(split_names) + 1
    split = split_module(gm, root, lambda n: stage_of[n])
    # parameters / buffers fetched at the top level move into the one stage that uses them (reference utils/trace.py:95-175)
    lift_single_use_params(split)
    # top-level plumbing: placeholders, call_module(submod_k), getitem, output
    producers: Dict[str, int] = {}       # value name -> producing stage (-1 for batch inputs)
    consumers: Dict[str, List[int]] = {}
    stage_inputs: Dict[int, List[str]] = {}
    stage_outputs: Dict[int, List[str]] = {}
    getitems: Dict[str, tuple] = {}
    out_struct = None

    def vname(n: fx.Node) -> str:
        return n.name

    for node in split.graph.nodes:
        if node.op == "placeholder":
            producers[vname(node)] = -1
        elif node.op == "call_module" and str(node.target).startswith("submod_"):
            k = int(str(node.target).split("_")[1])
            ins = [vname(a) for a in node.args if isinstance(a, fx.Node)]
            stage_inputs[k] = ins
            for a in ins:
                consumers.setdefault(a, []).append(k)
            producers[vname(node)] = k
            stage_outputs.setdefault(k, [])
        elif node.op == "call_function" and node.target.__name__ == "getitem":
            src, idx = node.args
            getitems[vname(node)] = (vname(src), idx)
            producers[vname(node)] = producers[vname(src)]
        elif node.op == "output":
            out_struct = node.args[0]

    def root_value(name):                 # resolve getitem chains to (submodule output name, index path)
        path = []
        while name in getitems:
            name, idx = getitems[name]
            path.append(idx)
        return name, tuple(reversed(path))

    # names that are used as *values* (after getitem resolution) by stages / the final output
    def collect_out_names(o, acc):
        if isinstance(o, fx.Node):
            acc.append(vname(o))
        elif isinstance(o, (list, tuple)):
            for x in o:
                collect_out_names(x, acc)
        elif isinstance(o, dict):
            for x in o.values():
                collect_out_names(x, acc)
    final_names: List[str] = []
    collect_out_names(out_struct, final_names)
    for n in final_names:
        consumers.setdefault(n, []).append(num_stages)     # pseudo-consumer after the last stage

    specs: List[StageSpec] = []
    for k in range(num_stages):
        sub = getattr(split, f"submod_{k}")
        ins = stage_inputs.get(k, [])
        load = [n for n in ins if producers.get(n) == -1]
        # live across boundary k-1 -> k: produced in a stage < k (not a batch input), consumed in a stage >= k
        recv = sorted(n for n, p in producers.items() if 0 <= p < k and any(c >= k for c in consumers.get(n, [])))
        send = sorted(n for n, p in producers.items() if 0 <= p <= k and any(c > k for c in consumers.get(n, []))) \
            if k < num_stages - 1 else []
        specs.append(StageSpec(k, _FxStage(sub, ins, k, getitems, producers), recv, load, send))
    last = specs[-1]
    last.output_names = final_names

    def rebuild(values: Dict[str, Any], struct=out_struct):
        def go(o):
            if isinstance(o, fx.Node):
                return values[vname(o)]
            if isinstance(o, tuple):
                return tuple(go(x) for x in o)
            if isinstance(o, list):
                return [go(x) for x in o]
            if isinstance(o, dict):
                return {kk: go(x) for kk, x in o.items()}
            return o
        return go(struct)
    last.rebuild_output = rebuild
    return specs


class _FxStage(nn.Module):
    """Runs one split GraphModule on named values and returns every value it defines (outputs + getitems)."""

    def __init__(self, sub: nn.Module, arg_names: List[str], index: int, getitems, producers):
        super().__init__()
        self.sub = sub
        self.arg_names = arg_names
        self.index = index
        self._getitems = {n: g for n, g in getitems.items() if producers.get(n) == index}
        self._my_outputs = [n for n, p in producers.items() if p == index and n not in getitems]

    def forward(self, **values):
        out = self.sub(*[values[n] for n in self.arg_names])
        res = dict(values)
        for n in self._my_outputs:
            res[n] = out
        pending = dict(self._getitems)
        while pending:
            progressed = False
            for n, (src, idx) in list(pending.items()):
                if src in res:
                    res[n] = res[src][idx]
                    del pending[n]
                    progressed = True
            if not progressed:
                break
        return res


# ------------------------------------------------------------------------------------------------------------
# entry point
# ------------------------------------------------------------------------------------------------------------
def build_stages(model: nn.Module, split_points: Sequence[Union[str, nn.Module]],
                 input_names: Optional[List[str]] = None) -> List[StageSpec]:
    names = names_of_split_points(model, split_points)
    if hasattr(model, "pipeline_stages"):
        return model.pipeline_stages(names)
    return trace_and_split(model, names, input_names)
