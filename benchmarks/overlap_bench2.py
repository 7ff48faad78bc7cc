"""Engine-like launch pattern: per iteration an all-gather is started on the side stream right before a qkv-sized GEMM
on the main stream; every GEMM is timed individually (CUDA events).  torchrun --nproc-per-node 2 ..."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    rank = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
    world = dist.get_world_size()
    dev = torch.device("cuda", rank)
    from torchacc_b200 import _native as nat
    from torchacc_b200.ops.linear import gemm
    from torchacc_b200.parallel.collectives import make_collectives
    coll = make_collectives(dist.group.WORLD, dev, True)
    nat.set_gemm_scheduler(os.environ.get("OVERLAP_SCHED", "dynamic") == "dynamic")
    N = int(os.environ.get("OVERLAP_N", "6144"))
    skew_us = int(os.environ.get("OVERLAP_SKEW_US", "0"))      # rank 1 starts each iteration late by this much
    n_layer = 218_112_000 // (8 * world) * (8 * world)
    shard = coll.alloc(n_layer // world, torch.bfloat16)
    fulls = [torch.empty(n_layer, dtype=torch.bfloat16, device=dev) for _ in range(2)]
    shard.normal_()
    x = torch.randn(8192, 4096, device=dev, dtype=torch.bfloat16)
    w = torch.randn(N, 4096, device=dev, dtype=torch.bfloat16) * 0.02
    y = torch.empty(8192, N, device=dev, dtype=torch.bfloat16)
    filler_a = torch.randn(8192, 8192, device=dev, dtype=torch.bfloat16)
    side = torch.cuda.Stream(dev, priority=-1)
    iters = 16

    def run(with_ag):
        evs = []
        dist.barrier()
        torch.cuda.synchronize()
        for i in range(iters):
            if skew_us and rank == 1:
                torch.cuda._sleep(int(skew_us * 1400))
            if with_ag:
                ev = torch.cuda.Event()
                ev.record()
                side.wait_event(ev)
                with torch.cuda.stream(side):
                    coll.all_gather(shard, fulls[i % 2])
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            gemm(x, w, out=y)
            b.record()
            evs.append((a, b))
            filler_a.mul_(1.0)            # ~0.1 ms of unrelated work between iterations
            if with_ag:
                torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        return [a.elapsed_time(b) for a, b in evs]

    run(True); run(False)
    alone = run(False)
    ovl = run(True)
    if rank == 0:
        print(json.dumps({"N": N, "sched": os.environ.get("OVERLAP_SCHED", "dynamic"), "skew_us": skew_us,
                          "tma_min": os.environ.get("TORCHACC_B200_COMM_TMA_MIN", "default"),
                          "gemm_alone_ms": [round(t, 3) for t in alone[4:]],
                          "gemm_with_ag_ms": [round(t, 3) for t in ovl[4:]]}))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
