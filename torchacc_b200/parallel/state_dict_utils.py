"""Sharded checkpoints: metadata, optimizer-state APIs, offline consolidate / reshard.

Capability parity with reference torchacc/dist/state_dict_utils.py:27-738 and dist/fsdp.py:243-578, on OUR shard
layout: every FSDP unit is one flat fp32 vector padded to ``128 * world`` (same padding rule as the reference,
state_dict_utils.py:355-357) and rank r owns ``flat[r*n:(r+1)*n]``.  ``shard_metadata`` records, per unit, the
ordered ``(name, shape, numel, offset)`` of the parameters inside the flat vector, so everything here is pure tensor
surgery that runs on CPU.

File recipe (same as the reference docs, docs/source/dist/fsdp.md:126-172): each rank saves
``{'model': model.sharded_state_dict(), 'shard_metadata': model.get_shard_metadata()}`` to
``rank{R}-of-{W}-model.pth`` and ``{'optimizer': optim.state_dict(), 'shard_metadata': ...}`` to
``rank{R}-of-{W}-optim.pth``; the CLI ``consolidate_and_reshard_fsdp_ckpts`` turns them into one full checkpoint or
into shards for a different world size.

Reference defects not reproduced: every shard file is validated (not just the first), the
CLI's ``--ckpt_type model|optimizer`` paths work, and optimizer consolidation does not need a side-car
``layer_info.pickle`` (the metadata travels inside every file).
"""
from __future__ import annotations

import glob
import os
from collections import OrderedDict
from concurrent.futures import ThreadPoolExecutor
from typing import Any, Dict, List, Optional, Tuple

import torch
import torch.distributed as dist

PAD_MULTIPLE = 128
_TENSOR_STATES = ("exp_avg", "exp_avg_sq", "momentum_buffer", "max_exp_avg_sq", "sum", "square_avg")


# ------------------------------------------------------------------------------------------------------------
# metadata
# ------------------------------------------------------------------------------------------------------------
def padded_numel(numel: int, world: int) -> int:
    m = PAD_MULTIPLE * world
    return max(m, (numel + m - 1) // m * m)


def get_shard_metadata(engine) -> Dict[str, Any]:
    units = []
    for u in engine.units:
        units.append({
            "prefix": u.prefix,
            "numel": u.numel,
            "padded": u.padded,
            "shard_numel": u.shard_numel,
            "params": [{"fqn": i.fqn, "shape": list(i.shape), "numel": i.numel, "offset": i.offset} for i in u.infos],
        })
    return {"world_size": engine.shard_world, "rank": engine.shard_rank, "pad_multiple": PAD_MULTIPLE,
            "units": units, "format": "torchacc_b200.flat.v1"}


def _full_name(unit_meta, p) -> str:
    return (unit_meta["prefix"] + "." if unit_meta["prefix"] else "") + p["fqn"]


def unflatten_params(flat: torch.Tensor, unit_meta) -> "OrderedDict[str, torch.Tensor]":
    out = OrderedDict()
    for p in unit_meta["params"]:
        out[_full_name(unit_meta, p)] = flat[p["offset"]:p["offset"] + p["numel"]].view(p["shape"]).clone()
    return out


def flatten_params(named: Dict[str, torch.Tensor], unit_meta, dtype=torch.float32) -> torch.Tensor:
    flat = torch.zeros(unit_meta["numel"], dtype=dtype)
    for p in unit_meta["params"]:
        flat[p["offset"]:p["offset"] + p["numel"]] = named[_full_name(unit_meta, p)].reshape(-1).to(dtype)
    return flat


def shard_flat(flat: torch.Tensor, world: int) -> List[torch.Tensor]:
    """Pad ``flat`` (unpadded length) to ``128*world`` and split into ``world`` equal shards."""
    padded = padded_numel(flat.numel(), world)
    buf = torch.zeros(padded, dtype=flat.dtype)
    buf[:flat.numel()] = flat
    return list(buf.chunk(world))


# ------------------------------------------------------------------------------------------------------------
# online: model + optimizer state (collective)
# ------------------------------------------------------------------------------------------------------------
def _gather_full(engine, shard: torch.Tensor) -> torch.Tensor:
    if engine.shard_world == 1:
        return shard
    full = torch.empty(shard.numel() * engine.shard_world, dtype=shard.dtype, device=shard.device)
    engine.shard_coll.all_gather(shard.contiguous(), full)
    return full


@torch.no_grad()
def full_model_state_dict(engine, rank0_only: bool = True, cpu_offload: bool = True) -> Dict[str, torch.Tensor]:
    meta = get_shard_metadata(engine)
    out = OrderedDict()
    keep = (not rank0_only) or _global_rank() == 0
    for u, um in zip(engine.units, meta["units"]):
        full = _gather_full(engine, u.flat_param.data)
        if keep:
            named = unflatten_params(full[:um["numel"]], um)
            for k, v in named.items():
                out[k] = v.cpu() if cpu_offload else v
    return out if keep else {}


def _global_rank() -> int:
    return dist.get_rank() if (dist.is_available() and dist.is_initialized()) else 0


def _flat_index(engine, optim) -> Dict[int, int]:
    """optimizer param index -> engine unit index."""
    by_id = {id(u.flat_param): i for i, u in enumerate(engine.units)}
    mapping, idx = {}, 0
    for g in optim.param_groups:
        for p in g["params"]:
            if id(p) in by_id:
                mapping[idx] = by_id[id(p)]
            idx += 1
    return mapping


def sharded_optim_state_dict(engine, optim) -> Dict[str, Any]:
    """This rank's optimizer state + the metadata needed to consolidate it (reference fsdp.py:243-289)."""
    return {"optimizer": optim.state_dict(), "shard_metadata": get_shard_metadata(engine)}


@torch.no_grad()
def full_optim_state_dict(engine, optim, rank0_only: bool = True, cpu_offload: bool = True) -> Dict[str, Any]:
    """Un-sharded, per-parameter optimizer state keyed by parameter name (reference fsdp.py:291-424)."""
    meta = get_shard_metadata(engine)
    sd = optim.state_dict()
    mapping = _flat_index(engine, optim)
    keep = (not rank0_only) or _global_rank() == 0
    state = OrderedDict()
    for idx, uidx in mapping.items():
        st = sd["state"].get(idx)
        if st is None:
            continue
        um = meta["units"][uidx]
        per_param: Dict[str, Dict[str, Any]] = {_full_name(um, p): {} for p in um["params"]}
        for key, val in st.items():
            if isinstance(val, torch.Tensor) and val.numel() == engine.units[uidx].shard_numel:
                full = _gather_full(engine, val.to(engine.device))
                if keep:
                    for name, t in unflatten_params(full[:um["numel"]], um).items():
                        per_param[name][key] = t.cpu() if cpu_offload else t
            elif keep:
                for name in per_param:
                    per_param[name][key] = val.clone() if isinstance(val, torch.Tensor) else val
        if keep:
            state.update(per_param)
    if not keep:
        return {}
    groups = []
    for g in sd["param_groups"]:
        g2 = {k: v for k, v in g.items() if k != "params"}
        names = []
        for idx in g["params"]:
            if idx in mapping:
                um = meta["units"][mapping[idx]]
                names += [_full_name(um, p) for p in um["params"]]
        g2["params"] = names
        groups.append(g2)
    return {"state": state, "param_groups": groups}


@torch.no_grad()
def optim_state_dict_to_load(engine, optim_state_dict: Dict[str, Any], rank0_only: bool = True) -> Dict[str, Any]:
    """Convert a sharded or full optimizer state dict into what ``optim.load_state_dict`` expects on THIS rank
    (reference fsdp.py:426-578).  With ``rank0_only`` a full dict only has to exist on rank 0."""
    if "optimizer" in optim_state_dict and "shard_metadata" in optim_state_dict:
        sm = optim_state_dict["shard_metadata"]
        if sm["world_size"] != engine.shard_world:
            raise ValueError(f"sharded optimizer state was saved with world size {sm['world_size']} but the current "
                             f"one is {engine.shard_world}; reshard it with consolidate_and_reshard_fsdp_ckpts")
        return optim_state_dict["optimizer"]
    full = optim_state_dict
    world, rank = engine.shard_world, engine.shard_rank
    if rank0_only and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        box = [full if _global_rank() == 0 else None]
        dist.broadcast_object_list(box, src=0)
        full = box[0]
    meta = get_shard_metadata(engine)
    state, groups = {}, []
    for uidx, um in enumerate(meta["units"]):
        names = [_full_name(um, p) for p in um["params"]]
        have = [n for n in names if n in full["state"]]
        if not have:
            continue
        keys = full["state"][have[0]].keys()
        st = {}
        for key in keys:
            sample = full["state"][have[0]][key]
            if isinstance(sample, torch.Tensor) and sample.dim() > 0 and key != "step":
                flat = flatten_params({n: full["state"][n][key] for n in names}, um)
                st[key] = shard_flat(flat, world)[rank].clone().to(engine.device)
            else:
                st[key] = sample.clone() if isinstance(sample, torch.Tensor) else sample
        state[uidx] = st
    for g in full["param_groups"]:
        g2 = {k: v for k, v in g.items() if k != "params"}
        g2["params"] = list(range(len(meta["units"])))
        groups.append(g2)
    if len(groups) != 1:
        # several groups: map each group's names back to unit indices
        name_to_unit = {}
        for uidx, um in enumerate(meta["units"]):
            for p in um["params"]:
                name_to_unit[_full_name(um, p)] = uidx
        for g2, g in zip(groups, full["param_groups"]):
            g2["params"] = sorted({name_to_unit[n] for n in g["params"] if n in name_to_unit})
    return {"state": state, "param_groups": groups}


# ------------------------------------------------------------------------------------------------------------
# offline: load / save / consolidate / reshard
# ------------------------------------------------------------------------------------------------------------
def load_checkpoints(ckpt_dir: str, ckpt_name_pattern: str) -> List[Dict[str, Any]]:
    """Load every shard file matching the glob pattern, validate ALL of them and return them ordered by rank."""
    paths = sorted(glob.glob(os.path.join(ckpt_dir, ckpt_name_pattern)))
    if not paths:
        raise FileNotFoundError(f"no checkpoint files match {os.path.join(ckpt_dir, ckpt_name_pattern)}")
    with ThreadPoolExecutor(max_workers=min(8, len(paths))) as ex:
        ckpts = list(ex.map(lambda p: torch.load(p, map_location="cpu", weights_only=False), paths))
    for p, c in zip(paths, ckpts):
        if "shard_metadata" not in c:
            raise ValueError(f"{p} has no 'shard_metadata'")
    ckpts.sort(key=lambda c: c["shard_metadata"]["rank"])
    world = ckpts[0]["shard_metadata"]["world_size"]
    if len(ckpts) != world:
        raise ValueError(f"expected {world} shard files, found {len(ckpts)}")
    for r, c in enumerate(ckpts):
        sm = c["shard_metadata"]
        if sm["rank"] != r or sm["world_size"] != world:
            raise ValueError(f"shard {r}: inconsistent metadata (rank={sm['rank']}, world={sm['world_size']})")
    return ckpts


def save_checkpoints(state_dicts: List[Dict[str, Any]], shard_metadatas: List[Dict[str, Any]], save_paths: List[str],
                     save_type: str = "model") -> None:
    """Write one file per (re)shard; ``save_type`` is 'model' or 'optimizer'."""
    def write(args):
        sd, sm, path = args
        os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
        torch.save({save_type: sd, "shard_metadata": sm}, path)
    with ThreadPoolExecutor(max_workers=min(8, max(len(save_paths), 1))) as ex:
        list(ex.map(write, zip(state_dicts, shard_metadatas, save_paths)))


def _engine_module(model):
    return model._inner_engine_module() if hasattr(model, "_inner_engine_module") else model


def _partition_coords() -> Dict[str, int]:
    """Where this rank sits on the axes that hold DIFFERENT parameters (tp slice, pp stage) or identical copies
    (replica = dp x sp) -- taken from the mesh ``accelerate()`` built; all zeros / ones without one."""
    out = {"tp": 0, "tp_num": 1, "pp": 0, "pp_num": 1, "replica": 0}
    try:
        from .. import get_global_context
        mesh = get_global_context().mesh
    except Exception:
        mesh = None
    if mesh is not None:
        out.update(tp=mesh.get_tp_rank(), tp_num=mesh.get_tp_num(), pp=mesh.get_pp_rank(), pp_num=mesh.get_pp_num())
        rep = mesh.get_group_ranks("replica") if mesh.get_replica_num() > 1 else None
        out["replica"] = rep.index(mesh.get_global_rank()) if rep else 0
    return out


def checkpoint_partition_dir(ckpt_dir: str, coords: Optional[Dict[str, int]] = None) -> str:
    """Directory holding the FSDP shard files of ONE (tp slice, pp stage): ``ckpt_dir`` itself for plain FSDP/DP,
    ``ckpt_dir/tp{t}-of-{T}_pp{p}-of-{P}`` under tensor / pipeline parallelism, where ranks with equal FSDP coordinates
    hold different parameters (round-1 advisor finding: they used to overwrite each other's files).  The
    consolidate / reshard CLI is run once per such directory."""
    c = coords or _partition_coords()
    if c["tp_num"] == 1 and c["pp_num"] == 1:
        return ckpt_dir
    return os.path.join(ckpt_dir, f"tp{c['tp']}-of-{c['tp_num']}_pp{c['pp']}-of-{c['pp_num']}")


def save_sharded_checkpoint(model, optimizer=None, ckpt_dir: str = ".", extra: Optional[Dict[str, Any]] = None) -> None:
    """The reference's documented FSDP recipe (docs/source/dist/fsdp.md:126-156) as one call: every rank writes
    ``rank{R}-of-{W}-model.pth`` (its fp32 flat shards + ``shard_metadata``) and, with an optimizer,
    ``rank{R}-of-{W}-optim.pth``; ``extra`` (step counter, LR-scheduler state, ...) goes into the model file.
    The files are what ``consolidate_and_reshard_fsdp_ckpts`` consumes (its default patterns match).

    R / W are FSDP shard coordinates.  Ranks that only replicate a shard (HSDP / DP / sequence-parallel peers) do not
    write; tensor- and pipeline-parallel partitions go to their own sub-directory (``checkpoint_partition_dir``)."""
    m = _engine_module(model)
    meta = m.get_shard_metadata()
    coords = _partition_coords()
    meta = dict(meta, partition={k: coords[k] for k in ("tp", "tp_num", "pp", "pp_num")})
    r, w = meta["rank"], meta["world_size"]
    d = checkpoint_partition_dir(ckpt_dir, coords)
    if coords["replica"] == 0:
        os.makedirs(d, exist_ok=True)
        payload = {"model": {k: v.detach().cpu() for k, v in m.sharded_state_dict().items()}, "shard_metadata": meta}
        if extra:
            payload["extra"] = extra
        torch.save(payload, os.path.join(d, f"rank{r}-of-{w}-model.pth"))
        if optimizer is not None:
            osd = m.sharded_optim_state_dict(optimizer)      # {'optimizer': state_dict, 'shard_metadata': meta}
            osd["shard_metadata"] = meta
            torch.save(osd, os.path.join(d, f"rank{r}-of-{w}-optim.pth"))
    if dist.is_available() and dist.is_initialized():
        dist.barrier()


def load_sharded_checkpoint(model, optimizer=None, ckpt_dir: str = ".") -> Dict[str, Any]:
    """Resume from ``save_sharded_checkpoint`` files written by the SAME world size (use the consolidate / reshard
    CLI for a different one; a mismatch raises like the reference, dist/fsdp.py:518-526).  Returns ``extra``."""
    m = _engine_module(model)
    meta = m.get_shard_metadata()
    r, w = meta["rank"], meta["world_size"]
    d = checkpoint_partition_dir(ckpt_dir)
    path = os.path.join(d, f"rank{r}-of-{w}-model.pth")
    if not os.path.exists(path):
        found = sorted(glob.glob(os.path.join(d, "rank*-of-*-model.pth")))
        raise FileNotFoundError(f"{path} not found (checkpoint world size differs? files: {found[:4]}); "
                                "reshard with `python -m torchacc_b200.utils.consolidate_and_reshard_ckpts`")
    ck = torch.load(path, map_location="cpu", weights_only=False)
    if ck["shard_metadata"]["world_size"] != w:
        raise ValueError(f"checkpoint was written by {ck['shard_metadata']['world_size']} ranks, this job has {w}")
    m.load_sharded_state_dict(ck["model"])
    if optimizer is not None:
        opath = os.path.join(d, f"rank{r}-of-{w}-optim.pth")
        if not os.path.exists(opath) and os.path.exists(opath.replace("-optim.pth", "-optimizer.pth")):
            opath = opath.replace("-optim.pth", "-optimizer.pth")      # files written by round-1 builds
        osd = torch.load(opath, map_location="cpu", weights_only=False)["optimizer"]
        optimizer.load_state_dict(osd)
    return ck.get("extra", {})


def consolidate_sharded_model_checkpoints(ckpts: List[Dict[str, Any]]) -> "OrderedDict[str, torch.Tensor]":
    meta = ckpts[0]["shard_metadata"]
    out = OrderedDict()
    for uidx, um in enumerate(meta["units"]):
        full = torch.cat([c["model"][f"flat_params.{uidx}"].reshape(-1) for c in ckpts])
        out.update(unflatten_params(full[:um["numel"]], um))
    return out


def consolidate_sharded_optimizer_checkpoints(ckpts: List[Dict[str, Any]]) -> Dict[str, Any]:
    meta = ckpts[0]["shard_metadata"]
    opt0 = ckpts[0]["optimizer"]
    state = OrderedDict()
    for idx, st0 in opt0["state"].items():
        um = meta["units"][idx]
        per_param = {_full_name(um, p): {} for p in um["params"]}
        for key, val in st0.items():
            if isinstance(val, torch.Tensor) and val.numel() == um["shard_numel"]:
                full = torch.cat([c["optimizer"]["state"][idx][key].reshape(-1) for c in ckpts])
                for name, t in unflatten_params(full[:um["numel"]], um).items():
                    per_param[name][key] = t
            else:
                for name in per_param:
                    per_param[name][key] = val.clone() if isinstance(val, torch.Tensor) else val
        state.update(per_param)
    groups = []
    for g in opt0["param_groups"]:
        g2 = {k: v for k, v in g.items() if k != "params"}
        g2["params"] = [_full_name(meta["units"][i], p) for i in g["params"] for p in meta["units"][i]["params"]]
        groups.append(g2)
    return {"state": state, "param_groups": groups}


def _reshard_meta(meta: Dict[str, Any], new_world: int, rank: int) -> Dict[str, Any]:
    units = []
    for um in meta["units"]:
        padded = padded_numel(um["numel"], new_world)
        u2 = dict(um)
        u2.update(padded=padded, shard_numel=padded // new_world)
        units.append(u2)
    m2 = dict(meta)
    m2.update(world_size=new_world, rank=rank, units=units)
    return m2


def reshard_model_dict(full: Dict[str, torch.Tensor], meta: Dict[str, Any], reshard_num: int
                       ) -> Tuple[List[Dict[str, torch.Tensor]], List[Dict[str, Any]]]:
    shards = [OrderedDict() for _ in range(reshard_num)]
    for uidx, um in enumerate(meta["units"]):
        flat = flatten_params(full, um)
        for r, piece in enumerate(shard_flat(flat, reshard_num)):
            shards[r][f"flat_params.{uidx}"] = piece.clone()
    return shards, [_reshard_meta(meta, reshard_num, r) for r in range(reshard_num)]


def reshard_optim_dict(full_optim: Dict[str, Any], meta: Dict[str, Any], reshard_num: int
                       ) -> Tuple[List[Dict[str, Any]], List[Dict[str, Any]]]:
    shards = [{"state": {}, "param_groups": []} for _ in range(reshard_num)]
    name_to_unit = {}
    for uidx, um in enumerate(meta["units"]):
        names = [_full_name(um, p) for p in um["params"]]
        for n in names:
            name_to_unit[n] = uidx
        have = [n for n in names if n in full_optim["state"]]
        if not have:
            continue
        for key, sample in full_optim["state"][have[0]].items():
            if isinstance(sample, torch.Tensor) and sample.dim() > 0 and key != "step":
                flat = flatten_params({n: full_optim["state"][n][key] for n in names}, um)
                for r, piece in enumerate(shard_flat(flat, reshard_num)):
                    shards[r]["state"].setdefault(uidx, {})[key] = piece.clone()
            else:
                for r in range(reshard_num):
                    shards[r]["state"].setdefault(uidx, {})[key] = sample.clone() if isinstance(sample, torch.Tensor) \
                        else sample
    for g in full_optim["param_groups"]:
        g2 = {k: v for k, v in g.items() if k != "params"}
        g2["params"] = sorted({name_to_unit[n] for n in g["params"] if n in name_to_unit})
        for r in range(reshard_num):
            shards[r]["param_groups"].append(dict(g2))
    return shards, [_reshard_meta(meta, reshard_num, r) for r in range(reshard_num)]


def consolidate_and_reshard_fsdp_model_dict(ckpt_dir: str, model_ckpt_name_pattern: str, save_dir: str = "",
                                            model_save_name_pattern: str = "", reshard_num: int = 1,
                                            save_model: bool = True):
    """Consolidate (``reshard_num == 1`` -> ``model_consolidated.pth``) or reshard the model shards."""
    ckpts = load_checkpoints(ckpt_dir, model_ckpt_name_pattern)
    meta = ckpts[0]["shard_metadata"]
    full = consolidate_sharded_model_checkpoints(ckpts)
    save_dir = save_dir or ckpt_dir
    if reshard_num == 1:
        if save_model:
            name = model_save_name_pattern or "model_consolidated.pth"
            os.makedirs(save_dir, exist_ok=True)
            torch.save({"model": full, "shard_metadata": _reshard_meta(meta, 1, 0)}, os.path.join(save_dir, name))
        return full, meta
    shards, metas = reshard_model_dict(full, meta, reshard_num)
    if save_model:
        pat = model_save_name_pattern or "rank*-of-*-model.pth"
        save_checkpoints(shards, metas, _expand(pat, save_dir, reshard_num), "model")
    return shards, metas


def consolidate_and_reshard_fsdp_optim_dict(ckpt_dir: str, optimizer_ckpt_name_pattern: str, save_dir: str = "",
                                            optimizer_save_name_pattern: str = "", reshard_num: int = 1,
                                            save_optimizer: bool = True):
    ckpts = load_checkpoints(ckpt_dir, optimizer_ckpt_name_pattern)
    meta = ckpts[0]["shard_metadata"]
    full = consolidate_sharded_optimizer_checkpoints(ckpts)
    save_dir = save_dir or ckpt_dir
    if reshard_num == 1:
        if save_optimizer:
            name = optimizer_save_name_pattern or "optimizer_consolidated.pth"
            os.makedirs(save_dir, exist_ok=True)
            torch.save({"optimizer": full, "shard_metadata": _reshard_meta(meta, 1, 0)}, os.path.join(save_dir, name))
        return full, meta
    shards, metas = reshard_optim_dict(full, meta, reshard_num)
    if save_optimizer:
        pat = optimizer_save_name_pattern or "rank*-of-*-optim.pth"
        save_checkpoints(shards, metas, _expand(pat, save_dir, reshard_num), "optimizer")
    return shards, metas


def consolidate_and_reshard_fsdp_checkpoint(ckpt_dir: str, model_ckpt_name_pattern: str,
                                            optimizer_ckpt_name_pattern: str, save_dir: str = "",
                                            model_save_name_pattern: str = "", optimizer_save_name_pattern: str = "",
                                            reshard_num: int = 1):
    consolidate_and_reshard_fsdp_model_dict(ckpt_dir, model_ckpt_name_pattern, save_dir, model_save_name_pattern,
                                            reshard_num)
    consolidate_and_reshard_fsdp_optim_dict(ckpt_dir, optimizer_ckpt_name_pattern, save_dir,
                                            optimizer_save_name_pattern, reshard_num)


def _expand(pattern: str, save_dir: str, world: int) -> List[str]:
    """``rank*-of-*-model.pth`` -> one path per rank (first '*' = rank, second '*' = world)."""
    paths = []
    for r in range(world):
        name, n = pattern, 0
        while "*" in name:
            name = name.replace("*", str(r if n == 0 else world), 1)
            n += 1
        paths.append(os.path.join(save_dir, name))
    return paths
