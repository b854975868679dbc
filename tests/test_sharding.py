"""Host logic of the multi-GPU path (frame sharding, regularizer row bands, pipelined allreduce
over contiguous gradient units) -- on CPU, world_size 2, gloo, with the oracle as each rank's
evaluator.  The sharded sum must equal the single-process objective."""
import importlib
import os
import socket
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sharding = importlib.import_module("super-resolution_b200.sharding")


def test_partitions_cover_everything_once():
    for n, world in [(16, 1), (16, 2), (9, 4), (5, 8), (64, 8)]:
        seen = sorted(k for r in range(world) for k in sharding.frame_shard(n, r, world))
        assert seen == list(range(n))
    for H, world in [(2048, 8), (28, 3), (7, 8), (1, 2)]:
        bands = [sharding.row_band(H, r, world) for r in range(world)]
        assert bands[0][0] == 0 and bands[-1][1] == H
        assert all(a[1] == b[0] for a, b in zip(bands, bands[1:]))
    for units, chunks in [(192, 4), (3, 8), (1, 4), (10, 3)]:
        cb = sharding.chunk_bounds(units, chunks)
        assert cb[0][0] == 0 and cb[-1][1] == units and len(cb) <= chunks
        assert all(a[1] == b[0] and a[1] > a[0] for a, b in zip(cb, cb[1:]))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


import pytest


@pytest.mark.parametrize("mixed", ["0", "1"])
def test_world2_gloo_sharded_objective_equals_full(tmp_path, oracle, mixed):
    """mixed = 1: one rank reports a single unit (cannot pipeline) -- all ranks must then agree on one
    collective per evaluation."""
    o = oracle
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=str(_free_port()), SRB_TEST_MIXED=mixed)
    outs = [str(tmp_path / ("rank%d.npy" % r)) for r in range(2)]
    procs = [subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "_sharding_worker.py"),
                               str(r), "2", outs[r]], env=env) for r in range(2)]
    for p in procs:
        assert p.wait(timeout=300) == 0
    got = [np.load(f) for f in outs]
    np.testing.assert_array_equal(got[0], got[1])       # every rank holds the same result
    # single-process reference on the same seeded inputs
    rng = np.random.default_rng(7)
    C, h, w, s, K, N = 2, 12, 10, 2, 3, 5
    psf = o.gaussian_psf(K, 1.0)
    shifts = rng.integers(-2, 3, size=(N, 2)).astype(np.float64)
    x = rng.random((C, h * s, w * s))
    lr = rng.random((N, C, h, w))
    wts = 0.5 + rng.random(x.shape)
    m = o.Model(s, psf, shifts)
    f, g = o.evaluate(m, x, o.upsample_observations(m, lr), o.REG_TV, 0.02, wts)
    n = x.size
    np.testing.assert_allclose(got[0][n], f, rtol=1e-13)
    assert np.linalg.norm(got[0][:n] - g.ravel()) <= 1e-13 * np.linalg.norm(g)


def test_world3_gloo_row_band_partition_equals_full(tmp_path, oracle):
    """Row-band partition (sharding.RowBandObjective): every rank starts with x valid on its own band only, the halo
    rows come from the neighbours, and the gradient bands -- which are never summed across ranks -- put together are
    the single-process gradient, bit for bit (same arithmetic on the same values)."""
    o = oracle
    world = 3
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=str(_free_port()), SRB_TEST_PARTITION="rows")
    outs = [str(tmp_path / ("rows%d.npy" % r)) for r in range(world)]
    procs = [subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "_sharding_worker.py"),
                               str(r), str(world), outs[r]], env=env) for r in range(world)]
    for p in procs:
        assert p.wait(timeout=300) == 0
    got = [np.load(f) for f in outs]
    rng = np.random.default_rng(7)
    C, h, w, s, K, N = 2, 18, 10, 2, 3, 5
    psf = o.gaussian_psf(K, 1.0)
    shifts = rng.integers(-1, 2, size=(N, 2)).astype(np.float64)
    x = rng.random((C, h * s, w * s))
    lr = rng.random((N, C, h, w))
    wts = 0.5 + rng.random(x.shape)
    m = o.Model(s, psf, shifts)
    f, g = o.evaluate(m, x, o.upsample_observations(m, lr), o.REG_TV, 0.02, wts)
    n = x.size
    total = sum(a[:n] for a in got)                       # bands are disjoint: the sum is their union
    np.testing.assert_array_equal(total, g.ravel())
    for a in got:                                         # the scalar exchange: every rank holds the same total
        np.testing.assert_allclose(a[n], float(np.sum(x ** 2)), rtol=1e-14)
    # band bookkeeping
    bands = [sharding.unit_band(12, r, world) for r in range(world)]
    assert bands == [(0, 4), (4, 8), (8, 12)]
    assert [sharding.unit_band(3, r, 8) for r in range(8)].count((0, 0)) >= 1
