#!/usr/bin/env python3
"""
Developer tool: what the box's PCIe path gives the host-array step (`step_host`), measured with
plain pinned-memory copies of the e2e workload's sizes (2^24 Lander3D envs: 268 MB of actions in,
755 MB of obs/reward/done out per step).  Prints one JSON line.

    gpurun -- python tools/pcie_probe.py
    gpurun --gpus 2 -- python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tools/pcie_probe.py        # every rank copies at once: the shared host path
"""
import os
import time
import json
import torch


def timed(fn, reps=10):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def concurrent(world, rank):
    """All ranks run the same pinned D2H / H2D copies at the same time; wall clock over a barrier."""
    import torch.distributed as dist
    dist.init_process_group('gloo')
    torch.cuda.set_device(rank)
    n = 1 << 24
    h_in = torch.zeros(n * 16, dtype=torch.uint8).pin_memory()
    h_out = torch.zeros(n * 45, dtype=torch.uint8).pin_memory()
    d_in = torch.zeros(n * 16, dtype=torch.uint8, device='cuda')
    d_out = torch.zeros(n * 45, dtype=torch.uint8, device='cuda')
    out = {'world': world}
    for name, fn, nbytes in (('d2h', lambda: h_out.copy_(d_out, non_blocking=True), h_out.numel()),
                             ('h2d', lambda: d_in.copy_(h_in, non_blocking=True), h_in.numel())):
        fn(); torch.cuda.synchronize(); dist.barrier()
        t0 = time.perf_counter()
        for _ in range(10):
            fn()
        torch.cuda.synchronize()
        mine = torch.tensor([time.perf_counter() - t0], dtype=torch.float64)
        dist.all_reduce(mine, op=dist.ReduceOp.MAX)
        out[name + '_all_ranks_at_once'] = {'ms': mine.item() * 100, 'aggregate_gbs': world * nbytes * 10 / mine.item() / 1e9,
                                            'per_gpu_gbs': nbytes * 10 / mine.item() / 1e9}
    if rank == 0:
        print(json.dumps(out))
    dist.destroy_process_group()


def main():
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if world > 1:
        return concurrent(world, int(os.environ.get('LOCAL_RANK', '0')))
    n = 1 << 24
    h_in = torch.zeros(n * 16, dtype=torch.uint8).pin_memory()
    h_out = torch.zeros(n * 45, dtype=torch.uint8).pin_memory()
    d_in = torch.zeros(n * 16, dtype=torch.uint8, device='cuda')
    d_out = torch.zeros(n * 45, dtype=torch.uint8, device='cuda')
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    out = {}
    ms = timed(lambda: d_in.copy_(h_in, non_blocking=True))
    out['h2d_alone'] = {'ms': ms, 'gbs': h_in.numel() / ms / 1e6}
    ms = timed(lambda: h_out.copy_(d_out, non_blocking=True))
    out['d2h_alone'] = {'ms': ms, 'gbs': h_out.numel() / ms / 1e6}

    def both():
        cur = torch.cuda.current_stream()
        s1.wait_stream(cur); s2.wait_stream(cur)
        with torch.cuda.stream(s1):
            d_in.copy_(h_in, non_blocking=True)
        with torch.cuda.stream(s2):
            h_out.copy_(d_out, non_blocking=True)
        cur.wait_stream(s1); cur.wait_stream(s2)
    ms = timed(both)
    out['both_directions'] = {'ms': ms, 'd2h_gbs': h_out.numel() / ms / 1e6, 'h2d_gbs': h_in.numel() / ms / 1e6}

    # the same bytes as 16 chunks (the step_host default), D2H only
    chunks_d = d_out.chunk(16)
    chunks_h = h_out.chunk(16)
    ms = timed(lambda: [h.copy_(d, non_blocking=True) for h, d in zip(chunks_h, chunks_d)])
    out['d2h_16_chunks'] = {'ms': ms, 'gbs': h_out.numel() / ms / 1e6}
    print(json.dumps(out))


if __name__ == '__main__':
    main()
