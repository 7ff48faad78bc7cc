"""Optimizer-state round trips of the FSDP engine -- the reference's matrix
(reference tests/distributed/test_fsdp_optim_state.py:153-343: full / sharded state dicts x rank0_only x flatten x
reshard 4 -> 2 x padding) on 4 gloo ranks: train 10 steps, take the optimizer state, load it into a NEW model (same or
halved shard world), take the state again and compare every tensor with ``torch.equal``."""
import pytest
import torch

from dist_utils import run_distributed


class Net(torch.nn.Module):
    def __init__(self, size):
        super().__init__()
        self.fc = torch.nn.ModuleList([torch.nn.Linear(size, size) for _ in range(5)])

    def forward(self, x):
        for f in self.fc:
            x = f(x)
        return x


def _init(size, fsdp_size, flatten):
    import torchacc_b200 as ta
    torch.manual_seed(0)
    model = Net(size)
    cfg = ta.Config()
    cfg.dist.fsdp.size = fsdp_size                 # < world: the remaining ranks replicate (HSDP), like a sub-group
    cfg.dist.fsdp.flatten_parameters = flatten     # API parity: this engine always shards one flat vector per unit
    cfg.dist.fsdp.wrap_layer_cls = {"Linear"}
    model = ta.accelerate(model, config=cfg)
    return model, torch.optim.AdamW(model.parameters(), lr=0.1)


def _train(model, optim, size, iters, update=True):
    optim.zero_grad()
    g = torch.Generator().manual_seed(7)
    for _ in range(iters):
        data = torch.rand(max(size, 8), size, generator=g)
        labels = torch.zeros(max(size, 8), dtype=torch.int64)
        loss = torch.nn.functional.nll_loss(torch.log_softmax(model(data), -1), labels)
        loss.backward()
        if update:
            optim.step()
            optim.zero_grad()


def _same(a, b):
    assert a["state"].keys() == b["state"].keys()
    for k in a["state"]:
        assert a["state"][k].keys() == b["state"][k].keys(), k
        for name, t1 in a["state"][k].items():
            t2 = b["state"][k][name]
            assert torch.equal(t1, t2) if isinstance(t1, torch.Tensor) else t1 == t2, (k, name)
    assert len(a["param_groups"]) == len(b["param_groups"])
    for g1, g2 in zip(a["param_groups"], b["param_groups"]):
        assert g1.keys() == g2.keys()
        for k in g1:
            assert g1[k] == g2[k], k


def _full_case(rank, world, size, rank0_only, flatten, new_world):
    import torch.distributed as dist
    m1, o1 = _init(size, world, flatten)
    _train(m1, o1, size, 10)
    osd1 = m1.full_optim_state_dict(o1, rank0_only=rank0_only)
    assert bool(osd1) == (not rank0_only or rank == 0)
    m2, o2 = _init(size, new_world, flatten)
    o2.load_state_dict(m2.optim_state_dict_to_load(osd1, rank0_only=rank0_only))
    _train(m2, o2, size, 1, update=False)          # forward + backward, no step: the state must be untouched
    osd2 = m2.full_optim_state_dict(o2, rank0_only=rank0_only)
    if osd1:
        _same(osd1, osd2)
    dist.barrier()


def _matrix_worker(rank, world, size):
    for rank0_only in (True, False):
        for flatten in (True, False):
            for new_world in (world, world // 2):
                _full_case(rank, world, size, rank0_only, flatten, new_world)


def test_full_optim_state_matrix_rank0_flatten_reshard():
    """8 reference cases: fsdp4 / fsdp4->2 x rank0_only / not x flatten / noflatten (model size 64)."""
    run_distributed(_matrix_worker, 4, args=(64,), timeout=600)


def _pad_worker(rank, world):
    # model_size = 4: every parameter is smaller than the 128 * world padding unit (reference "..._pad" cases)
    _full_case(rank, world, 4, False, True, world)
    _full_case(rank, world, 4, False, False, world)
    _full_case(rank, world, 4, True, True, world)


def test_full_optim_state_with_padding():
    run_distributed(_pad_worker, 4, timeout=600)


def _sharded_worker(rank, world):
    size = 64
    m1, o1 = _init(size, world, True)
    _train(m1, o1, size, 10)
    osd1 = m1.sharded_optim_state_dict(o1)
    assert osd1["shard_metadata"]["world_size"] == world and osd1["shard_metadata"]["rank"] == rank
    m2, o2 = _init(size, world, True)
    o2.load_state_dict(m2.optim_state_dict_to_load(osd1))
    _train(m2, o2, size, 1, update=False)
    osd2 = m2.sharded_optim_state_dict(o2)
    _same(osd1["optimizer"], osd2["optimizer"])
    # a sharded state dict of another world size is rejected with a pointer to the reshard tool
    m3, o3 = _init(size, world // 2, True)
    with pytest.raises(ValueError):
        m3.optim_state_dict_to_load(osd1)


def test_sharded_optim_state_roundtrip():
    run_distributed(_sharded_worker, 4, timeout=600)
