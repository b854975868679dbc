"""Worker of tests/test_sharding.py: one rank of a world_size-2 gloo group evaluating its frame
shard + regularizer row band with the CPU oracle behind `ShardedObjective` (the same host logic
bench.py runs over NCCL with the CUDA engine).  Writes its result to argv[3]."""
import importlib
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import sr_oracle as o  # noqa: E402

sharding = importlib.import_module("super-resolution_b200.sharding")


class OracleEvaluator:
    """Evaluator protocol on top of the CPU oracle: units are HR row groups of one channel."""

    def __init__(self, model, lr, reg_kind, lam, wts, rows, rows_per_unit):
        self.m, self.kind, self.lam, self.wts, self.rows = model, reg_kind, lam, wts, rows
        self.obs = o.upsample_observations(model, lr)
        self.C, self.H, self.W = wts.shape
        self.rpu = rows_per_unit
        self.tr = (self.H + rows_per_unit - 1) // rows_per_unit
        self.single = rows_per_unit >= self.H * self.C
        self._g = None

    def num_units(self):
        return 1 if self.single else self.C * self.tr

    def _first(self, u):
        if self.single:
            return u * self.C * self.H * self.W
        ch, t = divmod(u, self.tr)
        return ch * self.H * self.W + min(t * self.rpu, self.H) * self.W

    def unit_range(self, u0, u1):
        return self._first(u0), self._first(u1)

    def _partial(self, x):
        xs = x.numpy().reshape(self.C, self.H, self.W)
        f, g = o.evaluate(self.m, xs, self.obs, self.kind, 0.0, None)
        vals, parts = o.reg_apply_diff(self.kind, xs, self.lam * self.wts)
        r0, r1 = self.rows
        g[:, r0:r1] += parts[:, r0:r1]
        f += float(np.sum(self.lam * self.wts[:, r0:r1] * vals[:, r0:r1] ** 2))
        return f, g.reshape(-1)

    def eval_units(self, x, gc, u0, u1):
        if self._g is None:
            self._f, self._g = self._partial(x)
        b, e = self.unit_range(u0, u1)
        gc[b:e] = torch.from_numpy(self._g[b:e])

    def eval_finish(self, x, gc):
        gc[-1] = self._f
        self._g = None


class OracleRowEvaluator:
    """Evaluator of the row-band partition on the CPU oracle: units are groups of `rpu` HR rows of one channel; the
    oracle evaluates the whole objective on the rank's replica of x (NaN wherever the rank has no current data) and
    the rank keeps its own rows.  The cost split by tiles is the CUDA kernel's business (tests/test_gpu_multi.py);
    here the rank's cost share is a stand-in, sum of x^2 over its band, which only exercises the scalar exchange."""

    def __init__(self, model, lr, reg_kind, lam, wts, rpu):
        self.m, self.kind, self.lam, self.wts, self.rpu = model, reg_kind, lam, wts, rpu
        self.obs = o.upsample_observations(model, lr)
        self.C, self.H, self.W = wts.shape
        self.tr = (self.H + rpu - 1) // rpu

    def num_units(self):
        return self.C * self.tr

    def unit_range(self, u0, u1):
        def first(u):
            ch, t = divmod(u, self.tr)
            return ch * self.H * self.W + min(t * self.rpu, self.H) * self.W
        return first(u0), first(u1)

    def eval_unit_range(self, x, g, u0, u1, cost):
        b, e = self.unit_range(u0, u1)
        xs = x.numpy().reshape(self.C, self.H, self.W)
        with np.errstate(invalid="ignore"):
            _, grad = o.evaluate(self.m, xs, self.obs, self.kind, self.lam, self.wts)
        g[b:e] = torch.from_numpy(grad.reshape(-1)[b:e])
        cost[0] = float(np.sum(x.numpy()[b:e] ** 2))


def main_rows(rank, world, out):
    """Row-band partition: every rank holds every frame, x is current only on the rank's band; the halo comes
    from the neighbours (RowBandObjective.exchange_halo).  Output: [gradient band placed in a zero vector, cost]."""
    rng = np.random.default_rng(7)
    C, h, w, s, K, N = 2, 18, 10, 2, 3, 5
    psf = o.gaussian_psf(K, 1.0)
    shifts = rng.integers(-1, 2, size=(N, 2)).astype(np.float64)
    x = rng.random((C, h * s, w * s))
    lr = rng.random((N, C, h, w))
    wts = 0.5 + rng.random(x.shape)
    ev = OracleRowEvaluator(o.Model(s, psf, shifts), lr, o.REG_TV, 0.02, wts, rpu=6)
    n = x.size
    # reach of a gradient row into x: PSF twice + the shift twice (warp and its transpose) + 1 row of TV
    halo_rows = 2 * (K // 2) + 2 * 1 + 1
    obj = sharding.RowBandObjective(ev, n, w * s, halo_rows, dist=dist)
    x_local = torch.full((n,), float("nan"), dtype=torch.float64)
    x_local[obj.begin:obj.end] = torch.from_numpy(x.reshape(-1)[obj.begin:obj.end])
    g = torch.zeros(n, dtype=torch.float64)
    cost = torch.zeros(1, dtype=torch.float64)
    obj.evaluate(x_local, g, cost)
    assert not torch.isnan(g[obj.begin:obj.end]).any()
    np.save(out, np.concatenate([g.numpy(), cost.numpy()]))


def main():
    rank, world, out = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("gloo", rank=rank, world_size=world)
    if os.environ.get("SRB_TEST_PARTITION") == "rows":
        main_rows(rank, world, out)
        dist.destroy_process_group()
        return
    rng = np.random.default_rng(7)
    C, h, w, s, K, N = 2, 12, 10, 2, 3, 5
    psf = o.gaussian_psf(K, 1.0)
    shifts = rng.integers(-2, 3, size=(N, 2)).astype(np.float64)
    x = rng.random((C, h * s, w * s))
    lr = rng.random((N, C, h, w))
    wts = 0.5 + rng.random(x.shape)
    frames = sharding.frame_shard(N, rank, world)
    rows = sharding.row_band(h * s, rank, world)
    model = o.Model(s, psf, shifts[frames])
    # SRB_TEST_MIXED=1: rank 1 cannot pipeline (one unit = everything), rank 0 can
    rpu = 10 ** 6 if (os.environ.get("SRB_TEST_MIXED") == "1" and rank == 1) else 5
    ev = OracleEvaluator(model, lr[frames], o.REG_TV, 0.02, wts, rows, rows_per_unit=rpu)
    n = x.size
    gc = torch.zeros(n + 1, dtype=torch.float64)
    obj = sharding.ShardedObjective(ev, n, dist=dist, num_chunks=3)
    obj.evaluate(torch.from_numpy(x.reshape(-1).copy()), gc).wait()
    np.save(out, gc.numpy())
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
