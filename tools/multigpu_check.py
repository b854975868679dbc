#!/usr/bin/env python
"""Multi-GPU parity check (run under torchrun, one rank per GPU, NCCL):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 \
        --master-port 29511 tools/multigpu_check.py
Frames sharded round robin, regularizer split by row bands, pipelined allreduce: the result on
every rank must equal the single-GPU evaluation of the whole stack (rank 0 computes it on its own
GPU) up to fp64 reassociation (<= 1e-13 relative)."""
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


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ok = True
    for (C, h, w, s, K, N, reg) in [(3, 128, 160, 4, 7, 16, srb.REG_TV), (2, 96, 96, 2, 5, 8, srb.REG_BTV),
                                    (1, 100, 84, 3, 3, 5, srb.REG_TV)]:
        rng = np.random.default_rng(42)
        psf = wl.gaussian_psf(K, 1.5)
        shifts = rng.integers(-2, 3, size=(N, 2)).astype(np.float64) if reg == srb.REG_BTV else wl.default_shifts(N, s)
        x = rng.random((C, h * s, w * s))
        lr = rng.random((N, C, h, w))
        wts = 0.5 + rng.random(x.shape)
        n = x.size
        frames = sharding.frame_shard(N, rank, world)
        with srb.Engine((len(frames), C, h, w), s, psf, shifts[frames], device=local) as e:
            e.set_observations(lr[frames])
            e.set_regularizer(reg, 0.02)
            e.set_irls_weights(wts)
            e.set_regularizer_rows(*sharding.row_band(h * s, rank, world))
            stream = torch.cuda.ExternalStream(e.stream_handle())
            with torch.cuda.stream(stream):
                xd = torch.from_numpy(x.reshape(-1)).cuda()
                gc = torch.zeros(n + 1, dtype=torch.float64, device="cuda")
                obj = sharding.ShardedObjective(sharding.EngineEvaluator(e), n, dist=dist, num_chunks=4)
                obj.evaluate(xd, gc).wait()
                stream.synchronize()
            got = gc.cpu().numpy()
            units = e.num_units()[0]
            got_peer = None
            try:   # the fused reduce-scatter / peer-memory path, where the model qualifies
                with torch.cuda.stream(stream):
                    pobj = sharding.PeerObjective(e, n, dist, srb)
                    for _ in range(2):
                        pobj.evaluate(xd)
                    stream.synchronize()
                    got_peer = pobj.out[:n + 1].cpu().numpy()
                    pobj.close()
            except srb.SrbError as err:
                if rank == 0:
                    print("  peer path not applicable:", err, flush=True)
        if rank == 0:
            with srb.Engine(lr.shape, s, psf, shifts, device=local) as e:
                e.set_observations(lr)
                e.set_regularizer(reg, 0.02)
                e.set_irls_weights(wts)
                f, g = e.eval(x)
            rel = np.linalg.norm(got[:n] - g.ravel()) / np.linalg.norm(g)
            relf = abs(got[n] - f) / abs(f)
            good = rel <= 1e-13 and relf <= 1e-13
            if got_peer is not None:
                relp = np.linalg.norm(got_peer[:n] - g.ravel()) / np.linalg.norm(g)
                relpf = abs(got_peer[n] - f) / abs(f)
                print("  peer path: grad rel %.2e cost rel %.2e" % (relp, relpf), flush=True)
                good = good and relp <= 1e-13 and relpf <= 1e-13
            ok = ok and good
            print("world %d case C=%d %dx%d s=%d K=%d N=%d reg=%d units=%d: grad rel %.2e cost rel %.2e %s" %
                  (world, C, h * s, w * s, s, K, N, reg, units, rel, relf, "OK" if good else "FAIL"), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
