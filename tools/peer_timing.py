#!/usr/bin/env python
"""Phase timing of the multi-GPU peer path (torchrun, one rank per GPU): scatter phase (tile kernel
with peer stores + cost) vs gather phase (sum + peer stores + barriers), CUDA events on the
engine's stream."""
import importlib
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
srb = importlib.import_module("super-resolution_b200")
wl = importlib.import_module("super-resolution_b200.workloads")
sharding = importlib.import_module("super-resolution_b200.sharding")

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
local = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
cf = wl.CONFIGS[3]
H, W, C, s, N = cf["H"], cf["W"], cf["C"], cf["s"], cf["N"]
frames = sharding.frame_shard(N * world, rank, world)
shifts = wl.default_shifts(N * world, s)
rng = np.random.default_rng(rank)
lr = rng.random((N, C, H // s, W // s))
x = rng.random(C * H * W)
e = srb.Engine(lr.shape, s, wl.gaussian_psf(cf["K"], cf["sigma"]), shifts[frames], device=local)
e.set_observations(lr)
e.set_regularizer(srb.REG_TV, 0.01)
e.set_regularizer_rows(*sharding.row_band(H, rank, world))
stream = torch.cuda.ExternalStream(e.stream_handle())
with torch.cuda.stream(stream):
    xd = torch.from_numpy(x).cuda()
    dist.broadcast(xd, src=0)
    obj = sharding.PeerObjective(e, C * H * W, dist, srb)
    for _ in range(5):
        obj.evaluate(xd)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    ts, tg = [], []
    for _ in range(30):
        dist.barrier()
        ev[0].record(stream)
        e.peer_scatter_dev(xd)
        ev[1].record(stream)
        e.peer_gather_dev()
        ev[2].record(stream)
        stream.synchronize()
        ts.append(ev[0].elapsed_time(ev[1]))
        tg.append(ev[1].elapsed_time(ev[2]))
    print("rank %d world %d: scatter phase %.3f ms, gather phase (incl. both barriers) %.3f ms" %
          (rank, world, float(np.median(ts)), float(np.median(tg))), flush=True)
    # raw NVLink numbers between rank and its neighbour: one DMA copy, and a copy kernel (torch add)
    if world > 1:
        nbytes = 50 * 1024 * 1024
        peer_ptr = obj.slot_ptrs[(rank + 1) % world]
        dst = torch.as_tensor(sharding._DevArray(peer_ptr, nbytes // 8), device="cuda")
        src = torch.zeros(nbytes // 8, dtype=torch.float64, device="cuda")
        for name, fn in (("DMA copy (cudaMemcpyAsync)", lambda: dst.copy_(src)),
                         ("SM copy kernel (torch.add out=peer)", lambda: torch.add(src, 1.0, out=dst))):
            for _ in range(3):
                fn()
            dist.barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            for _ in range(10):
                fn()
            b.record(stream)
            stream.synchronize()
            if rank == 0:
                print("rank 0 -> peer, 50 MiB, %s: %.1f GB/s" % (name, nbytes * 10 / (a.elapsed_time(b) * 1e-3) / 1e9), flush=True)
            dist.barrier()
    obj.close()
dist.barrier()
dist.destroy_process_group()
e.close()
