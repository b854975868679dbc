#!/usr/bin/env python
"""bench.py -- throughput of the MAP objective's cost+gradient evaluation (the hot path).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config 2|3|4|5]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is ONE evaluation of ObjectiveFunction::ComputeAllTerms (data term over all LR frames + the
configuration's regularization term) on synthetic data of one of BASELINE.json's configurations; the
default is configuration 3 (2048x2048 HR RGB, 16 LR frames, 4x, 7x7 PSF, TV), the one the 60 %-of-HBM-
roofline target is quoted on.  Metric unit: HR px * frames * channels per second.

  value     device-resident evaluations (x, gradient, observations in HBM), CUDA-event timed
  e2e       the same evaluation through the C-ABI call a host solver makes: x copied from pinned host
            memory, gradient + cost copied back, every step.  N = 1: srb_eval.  N > 1: srb_multi_eval, ONE
            host thread (rank 0) driving all N GPUs the way the reference's single-threaded solver would
            (the other ranks wait on a CPU barrier and keep their GPUs free)
  solve     (N = 1) a whole inner solve behind one C-ABI call: srb_cg_minimize, 20 CG iterations of
            RunCGSolverAnalyticalDiff with every solver vector in HBM, pinned host x in / x out once
  roofline  algorithmic bytes (8*C*P*(3 + N/s^2)) / device time of the fused tile kernel
            (CUDA events recorded by the library around its launch, on the launching stream), against
            the measured HBM peak in MEASURED_PEAKS.json; `traffic` = DRAM bytes of one launch from
            the committed ncu capture (profiles/traffic.json)
  cpu_baseline  the CPU reference path (oracle/_ref: the reference's objective/regularizer sources +
            the C restatement of its OpenCV-backed data term) on this box's host cores, bounded sample,
            all cores and one thread

Multi-GPU (N > 1), one process per GPU.  The contract partition (BASELINE configuration 3): the
configuration's frames are SHARDED over the ranks -- strong scaling, `"scaling": "strong"` -- x replicated,
the regularizer split by HR row bands, one cross-rank sum of C*P+1 doubles per step (reduce-scatter by copy
engines over NVLink peer memory behind the tile kernel + gather kernel; NCCL allreduce where the peer path
does not apply).  The payload of that sum does not shrink with N while the per-rank kernel barely does (the
PSF passes are per evaluation, not per frame), so kernel-only strong scaling over the frame axis is < 1x by
construction; it is reported as measured.  `weak` in the same line: every rank holds the
configuration's full frame count (N x more frames in total), the arrangement round 1 reported.
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "MAP-solver gradient evaluations: HR px*frames*ch per second"
UNIT = "HRpx*frames*ch/s"
SOLVE_ITERS = 20
REG_OVERRIDE = None   # --reg


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=3, choices=[2, 3, 4, 5])
    ap.add_argument("--path", default="auto", choices=["auto", "reference_order", "fused"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample", type=int, default=1024, help="HR side of the CPU-baseline crop")
    ap.add_argument("--chunks", type=int, default=4, help="allreduce pipeline depth (N > 1, NCCL path)")
    ap.add_argument("--solve-iters", type=int, default=SOLVE_ITERS,
                    help="CG iterations of the timed device-resident solve (0 = skip)")
    ap.add_argument("--no-weak", action="store_true", help="N > 1: skip the weak-scaling measurement")
    ap.add_argument("--no-e2e", action="store_true", help="skip the end-to-end measurement (profiling runs)")
    ap.add_argument("--no-rows", action="store_true", help="N > 1: skip the row-band partition measurement")
    ap.add_argument("--reg", default=None, choices=["tv", "tv3d", "btv", "none"],
                    help="regularizer instead of the configuration's own (cfg4 is also quoted with 3-D TV)")
    return ap.parse_args()


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clock / throttle-reason sampling during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            parts = [t.strip() for t in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                smax.append(float(parts[2]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(smax)),
                       reasons=sorted(reasons), samples=len(sm))
        return out


# ---- the CPU reference path (oracle/: the checker, timed here as the baseline and nowhere shipped) ----
def _cpu_problem(cfg, side):
    from oracle import sr_oracle
    wl = importlib.import_module("super-resolution_b200.workloads")
    cf = wl.CONFIGS[cfg]
    side = min(side, cf["H"])
    w = wl.make(cfg, H=side, W=side, cheap=True)
    if REG_OVERRIDE is not None:
        w["reg_kind"] = wl.REG_KIND[REG_OVERRIDE]
    m = sr_oracle.Model(w["s"], w["psf"], w["shifts"])
    obs = sr_oracle.upsample_observations(m, w["lr"])
    return wl, w, m, obs, side


def cpu_reference_run(cfg, side, steps, warmup, threads, also_single_thread=False):
    """Times the CPU reference path on a bounded sample: a side x side HR crop of the workload (all
    channels, all frames).  Returns a dict for `cpu_baseline`."""
    from oracle import sr_oracle, sr_ref
    wl, w, m, obs, side = _cpu_problem(cfg, side)
    wts = np.ones_like(w["x0"])
    use_ref = sr_ref.available()
    fn = sr_ref.compute_all_terms if use_ref else sr_oracle.evaluate
    kw = dict(btv_range=w["btv_range"], btv_decay=w["btv_decay"])
    units = wl.work_units(side, side, w["C"], w["N"])

    def timed(nthreads, nsteps, nwarm):
        for _ in range(nwarm):
            fn(m, w["x0"], obs, w["reg_kind"], w["lam"], wts, threads=nthreads, **kw)
        t0 = time.perf_counter()
        for _ in range(nsteps):
            fn(m, w["x0"], obs, w["reg_kind"], w["lam"], wts, threads=nthreads, **kw)
        return (time.perf_counter() - t0) / nsteps

    dt = timed(threads, steps, warmup)
    out = {"value": units / dt, "unit": UNIT, "cores": threads, "kind": "port",
           "sample": "%s ComputeAllTerms on a %dx%dx%d HR crop, %d frames, %d thread(s), %d evals" %
                     ("oracle/_ref" if use_ref else "oracle port", side, side, w["C"], w["N"], threads, steps),
           "note": "reference-sources+port" if use_ref else "port", "seconds_per_eval": dt}
    if also_single_thread:
        # the reference's own loops are single-threaded; one evaluation on a quarter crop
        wl1, w1, m1, obs1, side1 = _cpu_problem(cfg, max(side // 2, 256))
        wts1 = np.ones_like(w1["x0"])
        t0 = time.perf_counter()
        fn(m1, w1["x0"], obs1, w1["reg_kind"], w1["lam"], wts1, threads=1, **kw)
        dt1 = time.perf_counter() - t0
        out["single_thread"] = {"value": wl1.work_units(side1, side1, w1["C"], w1["N"]) / dt1, "unit": UNIT, "cores": 1,
                                "sample": "one evaluation on a %dx%dx%d HR crop, 1 thread" % (side1, side1, w1["C"])}
    return out


def cpu_reference_solve(cfg, side, iters, threads):
    """The reference's own ALGLIB mincg (oracle/_ref, RunCGSolverAnalyticalDiff's configuration) on the CPU
    path for `iters` iterations, bounded sample.  Returns a dict for `solve`."""
    import ctypes as C
    from oracle import sr_oracle, sr_ref
    if not sr_ref.available():
        return None
    wl, w, m, obs, side = _cpu_problem(cfg, side)
    wts = np.ones_like(w["x0"])
    shape = w["x0"].shape
    kw = dict(btv_range=w["btv_range"], btv_decay=w["btv_decay"])
    evals = [0]
    FG = C.CFUNCTYPE(None, C.c_longlong, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_void_p)

    def fg(n, xp, fp, gp, user):
        x = np.ctypeslib.as_array(xp, shape=(n,)).reshape(shape)
        f, g = sr_ref.compute_all_terms(m, x, obs, w["reg_kind"], w["lam"], wts, threads=threads, **kw)
        fp[0] = f
        np.ctypeslib.as_array(gp, shape=(n,))[:] = g.reshape(-1)
        evals[0] += 1

    L = sr_ref.lib()
    L.ref_mincg.argtypes = [C.c_longlong, C.POINTER(C.c_double), C.c_double, C.c_double, C.c_double, C.c_int, FG,
                            C.c_void_p, C.POINTER(C.c_double)]
    x = np.ascontiguousarray(w["x0"], dtype=np.float64).reshape(-1).copy()
    rep = np.zeros(4)
    t0 = time.perf_counter()
    L.ref_mincg(x.size, x.ctypes.data_as(C.POINTER(C.c_double)), 0.0, 0.0, 0.0, int(iters), FG(fg), None,
                rep.ctypes.data_as(C.POINTER(C.c_double)))
    dt = time.perf_counter() - t0
    units = wl.work_units(side, side, w["C"], w["N"])
    return {"api": "alglib::mincgoptimize (RunCGSolverAnalyticalDiff's configuration) on the CPU path",
            "iterations": int(rep[0]), "evaluations": int(rep[1]), "termination_type": int(rep[2]), "seconds": dt,
            "value": units * int(rep[1]) / dt, "unit": UNIT, "final_cost": float(rep[3]),
            "sample": "%dx%dx%d HR crop, %d frames, %d thread(s) in the data term, ALGLIB single-threaded" %
                      (side, side, w["C"], w["N"], threads)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    steps = max(1, min(args.steps, 5))
    warm = max(1, min(args.warmup, 1))
    base = cpu_reference_run(args.config, args.cpu_sample, steps, warm, cores, also_single_thread=True)
    wl = importlib.import_module("super-resolution_b200.workloads")
    cf = wl.CONFIGS[args.config]
    line = {
        "impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": warm, "ms_per_step": base["seconds_per_eval"] * 1e3, "higher_is_better": True,
        "scaling": "strong" if args.gpus > 1 else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": cf["name"], "sample": base["sample"]},
        "cpu_baseline": base,
        "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    if args.solve_iters > 0:
        sol = cpu_reference_solve(args.config, min(args.cpu_sample, 512), args.solve_iters, cores)
        if sol is not None:
            line["solve"] = sol
    print(json.dumps(line))


def _watchdog(seconds, code=3):
    """A bench run must never hang a GPU box (a stuck collective at teardown would): hard exit after `seconds`."""
    import threading

    def fire():
        sys.stderr.write("bench.py: watchdog fired after %d s -- exiting with code %d\n" % (seconds, code))
        sys.stderr.flush()
        os._exit(code)
    t = threading.Timer(seconds, fire)
    t.daemon = True
    t.start()
    return t


def main():
    global REG_OVERRIDE
    args = parse_args()
    REG_OVERRIDE = args.reg
    _watchdog(int(os.environ.get("SRB_BENCH_WATCHDOG_S", "900")))
    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as dist
    srb = importlib.import_module("super-resolution_b200")
    wl = importlib.import_module("super-resolution_b200.workloads")
    sharding = importlib.import_module("super-resolution_b200.sharding")

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available() or srb.device_count() == 0:
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU fallback")
    torch.cuda.set_device(local_rank)
    cpu_group = None
    if world > 1:
        # NCCL kernels on a high-priority stream: they must be able to start while the tile kernel
        # of the following gradient slice still fills the SMs
        opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=os.environ.get("SRB_NCCL_PRIO", "1") == "1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank), pg_options=opts)
        cpu_group = dist.new_group(backend="gloo")   # host-side waits that leave the GPUs alone
    assert world == args.gpus or world == 1, "launch with torchrun --nproc-per-node == --gpus"

    cf = wl.CONFIGS[args.config]
    H, W, C, s = cf["H"], cf["W"], cf["C"], cf["s"]
    psf = wl.gaussian_psf(cf["K"], cf["sigma"])
    n = C * H * W
    dev = torch.device("cuda", local_rank)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def build(n_total):
        """Engine + workload of this rank's shard of an n_total-frame stack."""
        frames = sharding.frame_shard(n_total, rank, world)
        shifts_all = wl.default_shifts(n_total, s)
        assert frames, "more GPUs than frames"
        eng = srb.Engine((len(frames), C, H // s, W // s), s, psf, shifts_all[frames], device=local_rank)
        work = wl.make(args.config, forward=lambda k, plane: eng.forward(frames.index(k), plane), N=n_total, frames=frames)
        eng.set_observations(work["lr"])
        if args.reg is not None:
            work["reg_kind"] = wl.REG_KIND[args.reg]
        eng.set_regularizer(work["reg_kind"], work["lam"], work["btv_range"], work["btv_decay"])
        eng.set_regularizer_rows(*sharding.row_band(H, rank, world))
        eng.set_path({"auto": srb.PATH_AUTO, "reference_order": srb.PATH_REFERENCE_ORDER,
                      "fused": srb.PATH_FUSED}[args.path])
        return eng, work, frames

    def measure_resident(eng, x_dev, gc_dev, steps, warmup):
        """Device-resident steps of this rank's engine (+ the cross-rank sum for N > 1).
        Returns (ms per step, kernel launches, partition description, objective handles to close)."""
        stream = torch.cuda.ExternalStream(eng.stream_handle(), device=dev)
        objective = sharding.ShardedObjective(sharding.EngineEvaluator(eng), n, dist=dist if world > 1 else None,
                                              num_chunks=args.chunks)
        peer = None
        if world > 1 and os.environ.get("SRB_MULTI", "peer") == "peer":
            try:
                with torch.cuda.stream(stream):
                    peer = sharding.PeerObjective(eng, n, dist, srb)
            except Exception as err:   # all ranks raise together (sharding.PeerObjective)
                peer = None
                if rank == 0:
                    print("peer path not available (%s); using the NCCL allreduce path" % err, file=sys.stderr)

        def step():
            if peer is not None:
                peer.evaluate(x_dev)
            else:
                objective.evaluate(x_dev, gc_dev).wait()

        with torch.cuda.stream(stream):
            for _ in range(warmup):
                step()
            barrier()
            launches0 = eng.timing()["kernel_launches"]
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(steps):
                step()
            e1.record(stream)
            barrier()
            ms = e0.elapsed_time(e1) / steps
            launches = eng.timing()["kernel_launches"] - launches0
        part = ("single GPU" if world == 1 else
                "frame shard; band-pipelined tile kernel + copy-engine reduce-scatter over NVLink peer memory + "
                "sum/all-gather kernel (C*P+1 f64 per step)" if peer is not None else
                "frame shard + 1 NCCL allreduce(C*P+1 f64) per step, pipelined in %d slices" % args.chunks)
        return ms, launches, part, step, stream, peer, objective

    def reduce_max(*vals):
        t = torch.tensor(list(vals), dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(v) for v in t.cpu()]

    # ================= primary measurement: the configuration's frames (sharded when N > 1) =================
    n_frames = cf["N"]
    eng, work, frames = build(n_frames)
    units = wl.work_units(H, W, C, n_frames)
    alg_bytes = wl.algorithmic_bytes(H, W, C, len(frames), s, has_reg=True)
    x0 = np.ascontiguousarray(work["x0"]).reshape(-1)
    stream0 = torch.cuda.ExternalStream(eng.stream_handle(), device=dev)
    with torch.cuda.stream(stream0):
        x_dev = torch.from_numpy(x0).to(dev, non_blocking=False)
        gc_dev = torch.zeros(n + 1, dtype=torch.float64, device=dev)
    if world > 1:
        dist.broadcast(x_dev, src=0)        # replicas of the same estimate
    h_x = torch.from_numpy(x0.copy()).pin_memory()
    h_g = torch.empty(n + 1, dtype=torch.float64).pin_memory()

    warmup = max(args.warmup, 3)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    ms_step, launches, partition, step_resident, stream, peer, objective = measure_resident(eng, x_dev, gc_dev, args.steps, warmup)

    with torch.cuda.stream(stream):
        # ---- the dominant kernel alone: CUDA events recorded by the library around the tile kernel launch,
        #      on the launching stream (srb_set_profiling) ------------------------------------------------
        eng.set_profiling(True)
        kern_ms = []
        for _ in range(min(max(args.steps, 5), 50)):
            eng.eval_partial_dev(x_dev, gc_dev)
            stream.synchronize()
            kern_ms.append(eng.timing()["last_main_kernel_ms"])
        eng.set_profiling(False)
        kernel_ms = float(np.mean(kern_ms))
    barrier()

    # ---- end to end through the host-facing call ----------------------------------------------------------
    e_steps = max(3, min(args.steps, 20))
    cost = 0.0
    solve = None
    e2e_frames = None
    e2e_api = "srb_eval"
    if args.no_e2e:
        e2e_ms = float("nan")
    elif world == 1:
        for _ in range(3):
            eng.eval(h_x.numpy(), out=h_g.numpy()[:n])
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(e_steps):
            cost, _ = eng.eval(h_x.numpy(), out=h_g.numpy()[:n])
        e2e_ms = (time.perf_counter() - t0) * 1e3 / e_steps
    else:
        # ONE host thread drives all GPUs (srb_multi_eval); the other ranks wait on the CPU.  Two partitions:
        # rows (every device evaluates the whole objective on its HR row bands: no exchange, PCIe traffic and
        # kernel work per device both 1/N) is what a host solver would use and is reported as `e2e`; the
        # contract's frame shard is timed beside it as `e2e_frame_shard`.
        e2e_api = "srb_multi_eval (one host thread, %d GPUs, row-band partition)" % world
        e2e_ms = 0.0
        if rank == 0:
            shifts_all = wl.default_shifts(n_frames, s)
            with srb.Engine((n_frames, C, H // s, W // s), s, psf, shifts_all, device=local_rank) as gen:
                full = wl.make(args.config, forward=lambda k, plane: gen.forward(k, plane), N=n_frames)
            for part in (srb.PARTITION_FRAMES, srb.PARTITION_ROWS):
                with srb.MultiEngine((n_frames, C, H // s, W // s), s, psf, shifts_all, n_gpus=world, partition=part) as me:
                    me.set_observations(full["lr"])
                    me.set_regularizer(work["reg_kind"], full["lam"], full["btv_range"], full["btv_decay"])
                    for _ in range(3):
                        me.eval(h_x.numpy(), out=h_g.numpy()[:n])
                    t0 = time.perf_counter()
                    for _ in range(e_steps):
                        cost_p, _ = me.eval(h_x.numpy(), out=h_g.numpy()[:n])
                    ms_p = (time.perf_counter() - t0) * 1e3 / e_steps
                    if part == srb.PARTITION_ROWS and args.solve_iters > 0:
                        # the device-resident solve on all GPUs: every solver vector cut into the same row bands
                        try:
                            xs = h_x.clone().pin_memory()
                            me.cg_minimize_inplace(xs.numpy(), maxits=2)      # warm-up: workspace allocation
                            xs.copy_(h_x)
                            l0 = me.timing()["kernel_launches"]
                            t0 = time.perf_counter()
                            rep = me.cg_minimize_inplace(xs.numpy(), maxits=args.solve_iters)
                            dt = time.perf_counter() - t0
                            solve = {"api": "srb_multi_cg_minimize (RunCGSolverAnalyticalDiff, solver vectors in HBM cut into "
                                            "row bands over %d GPUs, one host thread)" % world,
                                     "iterations": rep["iterations"], "evaluations": rep["num_evaluations"],
                                     "termination_type": rep["termination_type"], "seconds": dt,
                                     "value": units * rep["num_evaluations"] / dt, "unit": UNIT,
                                     "ms_per_iteration": dt * 1e3 / max(rep["iterations"], 1),
                                     "h2d_bytes": n * 8, "d2h_bytes": n * 8,
                                     "gpu_launches": int(me.timing()["kernel_launches"] - l0), "final_cost": rep["final_cost"]}
                            del xs
                        except Exception as err:
                            solve = {"unavailable": str(err)[:200]}
                if part == srb.PARTITION_ROWS:
                    e2e_ms, cost = ms_p, cost_p
                else:
                    e2e_frames = {"value": units / (ms_p * 1e-3), "unit": UNIT, "ms_per_step": ms_p,
                                  "api": "srb_multi_eval (one host thread, %d GPUs, frame shard + NVLink "
                                         "all-gather / reduce-scatter)" % world,
                                  "h2d_bytes_per_step": n * 8, "d2h_bytes_per_step": (n + 1) * 8, "cost_check": cost_p}
            del full
        torch.cuda.set_device(local_rank)   # rank 0 drove every GPU from this thread
        dist.barrier(group=cpu_group)
    ms_step, e2e_ms, kernel_ms = reduce_max(ms_step, e2e_ms, kernel_ms)

    # ---- a whole inner solve behind one call (N = 1) -----------------------------------------------------
    if args.solve_iters > 0 and world == 1:
        xs = h_x.clone().pin_memory()
        eng.cg_minimize_inplace(xs.numpy(), maxits=2)          # warm-up: workspace allocation
        xs.copy_(h_x)
        l0 = eng.timing()["kernel_launches"]
        t0 = time.perf_counter()
        rep = eng.cg_minimize_inplace(xs.numpy(), maxits=args.solve_iters)
        dt = time.perf_counter() - t0
        solve = {"api": "srb_cg_minimize (RunCGSolverAnalyticalDiff with the solver vectors in HBM)",
                 "iterations": rep["iterations"], "evaluations": rep["num_evaluations"],
                 "termination_type": rep["termination_type"], "seconds": dt,
                 "value": units * rep["num_evaluations"] / dt, "unit": UNIT, "ms_per_iteration": dt * 1e3 / max(rep["iterations"], 1),
                 "h2d_bytes": n * 8, "d2h_bytes": n * 8, "gpu_launches": int(eng.timing()["kernel_launches"] - l0),
                 "final_cost": rep["final_cost"]}

    # keep the same step running until the clock sampler has seen >= ~1 s of load; the number of extra steps is
    # derived from the rank-reduced timings so that every rank runs the same count
    extra = int(max(0.0, 1.2 - ms_step * args.steps * 1e-3) / max(ms_step * 1e-3, 1e-6))
    with torch.cuda.stream(stream):
        for _ in range(min(extra, 20000)):
            step_resident()
        barrier()
    clocks = sampler.stop() if sampler else None
    path_name = {1: "reference_order", 2: "fused"}[eng.active_path]
    if path_name == "fused" and eng.zlayout_active:
        path_name = "fused_zlayout"
    if peer is not None:
        peer.close()
    del objective, peer

    # ================= weak scaling (N > 1): every rank holds the configuration's full frame count =========
    weak = None
    if world > 1 and not args.no_weak:
        eng.close()
        eng, work_w, frames_w = build(n_frames * world)
        ms_w, launches_w, part_w, step_w, stream_w, peer_w, obj_w = measure_resident(eng, x_dev, gc_dev, args.steps, warmup)
        (ms_w,) = reduce_max(ms_w)
        weak = {"scaling": "weak", "frames_per_gpu": n_frames, "frames_total": n_frames * world,
                "value": wl.work_units(H, W, C, n_frames * world) / (ms_w * 1e-3), "unit": UNIT, "ms_per_step": ms_w,
                "partition": part_w}
        if peer_w is not None:
            peer_w.close()
        del obj_w, peer_w

    # ================= row-band partition (N > 1): every rank holds every frame, no gradient exchange ========
    rows_block = None
    if world > 1 and not args.no_rows:
        eng.close()
        frames_all = list(range(n_frames))
        eng = srb.Engine((n_frames, C, H // s, W // s), s, psf, wl.default_shifts(n_frames, s), device=local_rank)
        work_r = wl.make(args.config, forward=lambda k, plane: eng.forward(k, plane), N=n_frames)
        eng.set_observations(work_r["lr"])
        eng.set_regularizer(work["reg_kind"], work_r["lam"], work_r["btv_range"], work_r["btv_decay"])
        try:
            stream_r = torch.cuda.ExternalStream(eng.stream_handle(), device=dev)
            with torch.cuda.stream(stream_r):
                halo = eng.halo_rows()
                robj = sharding.RowBandObjective(sharding.EngineEvaluator(eng), n, W, halo, dist=dist)
                cost_dev = torch.zeros(1, dtype=torch.float64, device=dev)
                for _ in range(warmup):
                    robj.evaluate(x_dev, gc_dev, cost_dev)
                barrier()
                # one step = halo exchange + this rank's kernels + scalar allreduce; captured once as a CUDA graph
                # (the launch sequence is identical every step) and replayed: the step is short enough for the CPU
                # side of five launches and three NCCL calls to show otherwise.  SRB_ROWS_GRAPH=0: eager.
                graph = None
                if os.environ.get("SRB_ROWS_GRAPH", "1") == "1":
                    try:
                        graph = torch.cuda.CUDAGraph()
                        with torch.cuda.graph(graph, stream=stream_r, capture_error_mode="thread_local"):
                            robj.evaluate(x_dev, gc_dev, cost_dev)
                    except Exception as gerr:
                        graph = None
                        if rank == 0:
                            print("row-band step not captured as a CUDA graph (%s): eager launches" % str(gerr)[:120], file=sys.stderr)
                ok = torch.tensor([1.0 if graph is not None else 0.0], dtype=torch.float64, device=dev)
                dist.all_reduce(ok, op=dist.ReduceOp.MIN)
                if float(ok.cpu()[0]) < 1.0:
                    graph = None
                run_step = graph.replay if graph is not None else (lambda: robj.evaluate(x_dev, gc_dev, cost_dev))
                for _ in range(3):
                    run_step()
                barrier()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream_r)
                for _ in range(args.steps):
                    run_step()
                e1.record(stream_r)
                barrier()
                (ms_r,) = reduce_max(e0.elapsed_time(e1) / args.steps)
                captured = graph is not None
                if graph is not None:      # the graph holds NCCL work: release it before the process group goes away
                    stream_r.synchronize()
                    graph.reset()
                    graph = None
                    run_step = None
                rows_block = {"scaling": "strong", "frames_per_gpu": n_frames, "frames_total": n_frames,
                              "value": units / (ms_r * 1e-3), "unit": UNIT, "ms_per_step": ms_r,
                              "cost_check": float(cost_dev.cpu()[0]),
                              "partition": "row bands of the HR image: every rank holds every frame and evaluates the whole "
                                           "objective on 1/%d of the (channel, tile row) units; per step %d halo rows of x "
                                           "to each neighbour (NCCL send/recv) + a scalar allreduce of the cost; no gradient "
                                           "exchange; step %s" % (world, halo, "replayed as one CUDA graph" if captured else "launched eagerly")}
        except Exception as err:    # e.g. a model with a border band (cfg4 / cfg5): unit ranges do not apply
            rows_block = {"unavailable": str(err)[:200]}
        del work_r

    if rank == 0:
        peak, peak_src = measured_peaks()
        achieved = alg_bytes / (kernel_ms * 1e-3) / 1e9
        kernel_names = {"fused": "k_tile (fused tile kernel), one launch per evaluation",
                        "fused_zlayout": "k_tile_zt (fused tile kernel, observations in the transposed Z layout), "
                                         "one launch per evaluation"}
        line = {
            "metric": METRIC, "value": units / (ms_step * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "strong" if world > 1 else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": cf["name"] + ("" if args.reg is None else " [regularizer: %s]" % args.reg) + ("" if world == 1 else " -- its %d frames sharded over %d GPUs" % (n_frames, world)),
                       "frames_per_gpu": len(frames), "frames_total": n_frames,
                       "partition": partition, "kernel_path": path_name,
                       "l2": "inputs larger than L2 (%.0f MB touched per step vs 126 MB L2)" % (alg_bytes / 1e6),
                       "cost_check": cost},
            "e2e": {"value": units / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms, "api": e2e_api,
                    "h2d_bytes_per_step": n * 8, "d2h_bytes_per_step": (n + 1) * 8},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                         "kernel_ms": kernel_ms, "algorithmic_bytes": alg_bytes,
                         "kernel": kernel_names.get(path_name, "reference-order kernels")},
        }
        prof = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(prof):
            try:
                line["roofline"]["traffic"] = json.load(open(prof)).get("cfg%d_%s" % (args.config, path_name))
            except Exception:
                pass
        plan = srb.plan((len(frames), C, H // s, W // s), s, psf, work["shifts"])
        if path_name == "fused_zlayout" and plan["zt_frames"] > 1:
            # frames with equal shifts are averaged once at upload (srb_kernels_tilez.cuh): the kernel then reads one
            # observation per HR pixel instead of N / s^2.  `achieved` stays SURVEY 8d's algorithmic bytes (what the
            # reference's algorithm has to touch); the bytes the kernel really moves are stated beside it.
            # (the weights are read by the tile kernel only when the regularizer is evaluated inside it: 2-D TV)
            moved = wl.algorithmic_bytes(H, W, C, len(frames) // plan["zt_frames"], s,
                                         has_reg=work["reg_kind"] == wl.REG_KIND["tv"])
            line["roofline"].update(merged_frames_per_phase=plan["zt_frames"], bytes_moved_model=moved,
                                    achieved_moved=moved / (kernel_ms * 1e-3) / 1e9,
                                    frac_moved=moved / (kernel_ms * 1e-3) / 1e9 / peak)
        if solve is not None:
            line["solve"] = solve
        if e2e_frames is not None:
            line["e2e_frame_shard"] = e2e_frames
        if weak is not None:
            line["weak"] = weak
        if rows_block is not None:
            line["row_bands"] = rows_block
        if not args.no_cpu_baseline and world == 1:
            cores = os.cpu_count() or 1
            line["cpu_baseline"] = cpu_reference_run(args.config, args.cpu_sample, 3, 1, cores, also_single_thread=True)
        print(json.dumps(line), flush=True)
    _watchdog(60, code=0)   # the line is out: nothing below (teardown of streams / process group) may keep the box busy
    # release every tensor that lives on the engine's stream before the stream goes away
    del x_dev, gc_dev, h_x, h_g
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    eng.close()


if __name__ == "__main__":
    main()
