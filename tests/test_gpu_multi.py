"""GPU tests of the single-process multi-GPU form (srb_multi_*, include/srb200.h): one host thread drives
G devices; frames sharded in contiguous blocks (objective_data_term.cpp:104-114 is the loop being split),
regularizer by row bands, gradient reduce-scattered over NVLink peer memory.  The G-device result must
equal the oracle's full objective and the single-device evaluation up to fp64 re-association (SURVEY 8e:
<= 1e-13 relative).  Cases that need more devices than the box has run with several contexts per GPU
(SRB_MULTI_SHARE_DEVICES=1)."""
from importlib import import_module

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
wl = import_module("super-resolution_b200.workloads")
REL = 1e-12


@pytest.fixture(scope="module")
def srb():
    import srb200
    assert srb200.device_count() > 0, "no CUDA device: the product has no CPU fallback"
    return srb200


def rel_l2(a, b):
    return float(np.linalg.norm((a - b).ravel()) / max(np.linalg.norm(b.ravel()), 1e-300))


def _case(N, s, K, sigma, C, h, w, seed, frac=False):
    rng = np.random.default_rng(seed)
    psf = wl.gaussian_psf(K, sigma)
    shifts = wl.default_shifts(N, s)
    if frac:
        shifts = shifts + rng.uniform(-0.4, 0.4, size=shifts.shape)
    lr = rng.random((N, C, h, w))
    x = rng.random((C, h * s, w * s))
    wts = rng.uniform(0.5, 2.0, size=x.shape)
    return psf, shifts, lr, x, wts


CASES = [
    # N, s, K, sigma, C, h, w, reg, frac
    (16, 4, 7, 1.5, 2, 48, 80, "tv", False),    # cfg3's model: Z layout with holes on every shard
    (5, 4, 7, 1.5, 1, 48, 80, "tv", False),     # fewer frames than devices at G = 8: empty shards
    (9, 4, 5, 1.2, 1, 40, 64, "btv", False),    # cfg2's model: BTV is not fused -> unpipelined exchange
    (8, 2, 5, 1.0, 3, 40, 64, "tv3d", False),   # cfg4's model: two frames per phase, 3-D TV
    (6, 2, 3, 0.8, 1, 64, 96, "none", True),    # fractional shifts, no regularizer
    (16, 4, 7, 1.5, 3, 72, 80, "btv", False),   # no border band + BTV (tiled kernel): pipelined in both partitions
    (32, 4, 5, 1.2, 2, 64, 64, "tv", False),    # two frames per phase (merged at upload in the rows partition)
]


@pytest.mark.parametrize("partition", ["frames", "rows"])
@pytest.mark.parametrize("G", [1, 2, 4, 8])
@pytest.mark.parametrize("N,s,K,sigma,C,h,w,reg,frac", CASES)
def test_multi_eval_matches_oracle_and_single_device(srb, oracle, monkeypatch, G, N, s, K, sigma, C, h, w, reg, frac, partition):
    """partition = rows: every device holds every frame and evaluates the whole objective on its HR row bands
    (no exchange); configurations that cannot be cut into row bands are evaluated by device 0 alone.
    On a box with fewer than G GPUs the G contexts are spread round-robin over the GPUs there are
    (SRB_MULTI_SHARE_DEVICES=1: own context, streams and buffers each) -- the same multi-device code, kernels
    reading their peers' buffers through plain device pointers instead of NVLink mappings."""
    devices = None
    if srb.device_count() < G:
        monkeypatch.setenv("SRB_MULTI_SHARE_DEVICES", "1")
        devices = [i % srb.device_count() for i in range(G)]
    part = srb.PARTITION_ROWS if partition == "rows" else srb.PARTITION_FRAMES
    psf, shifts, lr, x, wts = _case(N, s, K, sigma, C, h, w, seed=100 + N + K, frac=frac)
    kind = {"tv": srb.REG_TV, "tv3d": srb.REG_TV3D, "btv": srb.REG_BTV, "none": srb.REG_NONE}[reg]
    okind = {"tv": oracle.REG_TV, "tv3d": oracle.REG_TV3D, "btv": oracle.REG_BTV, "none": oracle.REG_TV}[reg]
    lam = 0.0 if reg == "none" else 0.01
    m = oracle.Model(s, psf, shifts)
    obs = oracle.upsample_observations(m, lr)
    cost_ref, g_ref = oracle.evaluate(m, x, obs, okind, lam, wts if lam > 0 else None)
    with srb.MultiEngine(lr.shape, s, psf, shifts, n_gpus=G, devices=devices, partition=part) as me, \
            srb.Engine(lr.shape, s, psf, shifts) as e1:
        for e in (me, e1):
            e.set_observations(lr)
            e.set_regularizer(kind, lam)
            if lam > 0:
                e.set_irls_weights(wts)
        c1, g1 = e1.eval(x)
        for rep in range(2):   # the second evaluation re-uses the slots of the first
            cm, gm = me.eval(x)
            assert abs(cm - cost_ref) <= REL * abs(cost_ref)
            assert rel_l2(gm, g_ref) <= REL
            assert abs(cm - c1) <= 1e-13 * abs(c1)
            assert rel_l2(gm, g1) <= 1e-13
        c_only, none = me.eval(x, want_grad=False)
        assert none is None and abs(c_only - cost_ref) <= REL * abs(cost_ref)
        # a channel sub-range (IRLSMapSolver's split_channels)
        if C > 1 and reg != "tv3d":
            me.set_channel_range(1, C)
            e1.set_channel_range(1, C)
            if lam > 0:
                me.set_irls_weights(wts[1:])
                e1.set_irls_weights(wts[1:])
            cm, gm = me.eval(x[1:])
            c1, g1 = e1.eval(x[1:])
            assert abs(cm - c1) <= 1e-13 * abs(c1) and rel_l2(gm, g1) <= 1e-13


def test_multi_reweight_replicates_weights(srb, oracle):
    G = min(srb.device_count(), 2)
    psf, shifts, lr, x, wts = _case(4, 2, 3, 0.8, 2, 32, 64, seed=9)
    with srb.MultiEngine(lr.shape, 2, psf, shifts, n_gpus=G) as me:
        me.set_observations(lr)
        me.set_regularizer(srb.REG_TV, 0.01)
        w = me.reweight(x)
        assert np.array_equal(w, oracle.reweight(oracle.REG_TV, x))
        m = oracle.Model(2, psf, shifts)
        obs = oracle.upsample_observations(m, lr)
        cost_ref, g_ref = oracle.evaluate(m, x, obs, oracle.REG_TV, 0.01, w)
        cm, gm = me.eval(x)
        assert abs(cm - cost_ref) <= REL * abs(cost_ref) and rel_l2(gm, g_ref) <= REL


def test_multi_rejects_bad_arguments(srb):
    psf = wl.gaussian_psf(3, 0.8)
    with pytest.raises(srb.SrbError):
        srb.MultiEngine((4, 1, 16, 16), 2, psf, wl.default_shifts(4, 2), n_gpus=9)
    with pytest.raises(srb.SrbError):
        srb.MultiEngine((4, 1, 16, 16), 2, psf, wl.default_shifts(4, 2), n_gpus=2, devices=[0, 0])
    with pytest.raises(srb.SrbError):
        srb.MultiEngine((0, 1, 16, 16), 2, psf, None, n_gpus=1)


@pytest.mark.parametrize("partition", ["frames", "rows"])
def test_reference_solver_drives_several_gpus_through_the_adapter(srb, oracle, ref, partition):
    """The reference's IRLS + ALGLIB loop (oracle/_ref, unmodified ALGLIB and RunCGSolverAnalyticalDiff) on top of
    CudaMultiObjectiveTerm (include/srb200_adapters.hpp -> srb_multi_eval): one host thread, every GPU of the box
    (at most 8).  Same solve as through CudaObjectiveTerm on one device: the gradients differ by fp64
    re-association of the cross-device sum only, and a tie-free image keeps the solver from amplifying that."""
    G = min(srb.device_count(), 8)
    s, K = 4, 7
    rng = np.random.default_rng(41)
    psf, shifts = wl.gaussian_psf(K, 1.5), wl.default_shifts(16, s)
    truth = wl.box_smooth(rng.random((2, 192, 320)))
    m = oracle.Model(s, psf, shifts)
    lr = np.stack([np.stack([oracle.forward(m, k, truth[c]) for c in range(2)]) for k in range(16)])
    lr += 0.004 * rng.standard_normal(lr.shape)
    x0 = wl.bilinear_upsample(lr[0], s)
    opt = ref.default_options()
    opt.max_num_solver_iterations = 25
    opt.max_num_irls_iterations = 2
    part = srb.PARTITION_ROWS if partition == "rows" else srb.PARTITION_FRAMES
    with srb.Engine(lr.shape, s, psf, shifts) as e1, \
            srb.MultiEngine(lr.shape, s, psf, shifts, n_gpus=G, partition=part) as me:
        for e in (e1, me):
            e.set_observations(lr)
            e.set_regularizer(srb.REG_TV, 0.01)
        one, st1 = ref.solve_fused(e1, x0, True, 0.01, options=opt)
        many, stm = ref.solve_fused_multi(me, x0, True, 0.01, options=opt)
    assert stm.num_data_term_evals == st1.num_data_term_evals
    assert rel_l2(many, one) <= 1e-9
    assert rel_l2(many, truth) < rel_l2(x0, truth)


@pytest.mark.parametrize("K,reg", [(7, "tv"), (7, "btv"), (9, "tv"), (9, "btv"), (5, "none")])
def test_unit_ranges_of_the_whole_objective(srb, oracle, K, reg):
    """srb_eval_unit_range_dev (the row-band partition with one process per GPU): a range of (channel, tile row)
    units evaluated from an estimate that is valid ONLY on the range's rows plus srb_halo_rows() rows either side
    gives exactly the gradient rows of the full evaluation, and the range costs add up to the full cost.  K = 9 and
    K = 5 models have a border band of special samples (bottom / right image edge), which is cut by rows too."""
    import torch
    sharding = import_module("super-resolution_b200.sharding")
    s = 4
    N = 32 if K == 5 else 16                      # K = 5: two frames per phase, merged at upload
    psf, shifts, lr, x, wts = _case(N, s, K, 1.5, 2, 72, 80, seed=5 + K)
    kind = {"tv": srb.REG_TV, "btv": srb.REG_BTV, "none": srb.REG_NONE}[reg]
    with srb.Engine(lr.shape, s, psf, shifts) as e:
        e.set_observations(lr)
        e.set_regularizer(kind, 0.01)
        if reg != "none":
            e.set_irls_weights(wts)
        f_full, g_full = e.eval(x)
        p = srb.plan(lr.shape, s, psf, shifts)
        has_band = p["band_hi_r"] < 72 or p["band_lo_r"] > 0
        assert has_band == (K != 7)
        nu = e.total_units()
        assert nu == 2 * 9
        ev = sharding.EngineEvaluator(e)
        halo = e.halo_rows() * e.W
        stream = torch.cuda.ExternalStream(e.stream_handle())
        with torch.cuda.stream(stream):
            g_dev = torch.zeros(x.size, dtype=torch.float64, device="cuda")
            cost_dev = torch.zeros(1, dtype=torch.float64, device="cuda")
            for world in (1, 3, 4):
                total = 0.0
                g_dev.zero_()
                for r in range(world):
                    u0, u1 = sharding.unit_band(nu, r, world)
                    b, en = ev.row_unit_range(u0, u1)
                    xl = np.full(x.size, np.nan)
                    lo, hi = max(b - halo, 0), min(en + halo, x.size)
                    xl[lo:hi] = x.reshape(-1)[lo:hi]
                    x_dev = torch.from_numpy(xl).cuda()
                    e.eval_unit_range_dev(x_dev, g_dev, u0, u1, cost_dev)
                    stream.synchronize()
                    total += float(cost_dev.cpu()[0])
                got = g_dev.cpu().numpy().reshape(x.shape)
                if has_band:     # the band adds into the tile kernel's rows: same values, same order
                    np.testing.assert_allclose(got, g_full, rtol=0, atol=1e-13 * np.abs(g_full).max())
                else:
                    np.testing.assert_array_equal(got, g_full)
                assert abs(total - f_full) <= 1e-13 * abs(f_full), (reg, world)
