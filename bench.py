#!/usr/bin/env python
"""bench.py -- throughput of the MAP objective's cost+gradient evaluation (the hot path).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config 3]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is ONE evaluation of ObjectiveFunction::ComputeAllTerms (data term over all LR frames +
TV regularization term) on synthetic data of BASELINE.json's configuration 3 (2048x2048 HR RGB,
16 LR frames, 4x, 7x7 PSF, TV) -- the configuration the 60 %-of-HBM-roofline target is quoted on.
Metric unit: HR px * frames * channels per second.

  value     device-resident evaluations (x, gradient, observations in HBM), CUDA-event timed
  e2e       the same evaluation through the C-ABI call a host solver makes (srb_eval): x copied
            from pinned host memory, gradient + cost copied back, every step
  roofline  algorithmic bytes (8*C*P*(3 + N/s^2)) / device time of the fused tile kernel
            (CUDA events recorded by the library around its launch, on the launching stream), against
            the measured HBM peak in MEASURED_PEAKS.json; `traffic` = DRAM bytes of one launch from
            the committed ncu capture (profiles/traffic.json)
  cpu_baseline  the CPU reference path (oracle/_ref: the reference's objective/regularizer sources +
            the C restatement of its OpenCV-backed data term) on this box's host cores, bounded sample

Multi-GPU (N > 1): frames are sharded over ranks (weak scaling: every rank holds `N_frames` frames
of a N*N_frames stack), x is replicated, the regularizer is split by HR row bands, and ONE NCCL
allreduce over C*P+1 doubles per step yields gradient and cost everywhere; the allreduce is cut
into contiguous slices that overlap the tile kernel's work on the following rows (sharding.py).
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


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=3)
    ap.add_argument("--path", default="auto", choices=["auto", "reference_order", "fused"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample", type=int, default=1024, help="HR side of the CPU-baseline crop")
    ap.add_argument("--chunks", type=int, default=4, help="allreduce pipeline depth (N > 1)")
    ap.add_argument("--solve-iters", type=int, default=0,
                    help="N = 1 only, off by default: also time one device-resident CG solve of this many "
                         "iterations through srb_cg_minimize (host x in, host x out) and add it as \"solve\"")
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


def cpu_reference_run(cfg, side, steps, warmup, threads):
    """Times the CPU reference path on a bounded sample: a side x side HR crop of the workload
    (all channels, all frames of one rank).  Returns (units_per_s, seconds_per_eval, description)."""
    from oracle import sr_oracle, sr_ref
    wl = importlib.import_module("super-resolution_b200.workloads")
    cf = wl.CONFIGS[cfg]
    side = min(side, cf["H"])
    w = wl.make(cfg, H=side, W=side, cheap=True)
    m = sr_oracle.Model(w["s"], w["psf"], w["shifts"])
    obs = sr_oracle.upsample_observations(m, w["lr"])
    wts = np.ones_like(w["x0"])
    use_ref = sr_ref.available()
    fn = sr_ref.compute_all_terms if use_ref else sr_oracle.evaluate
    kw = dict(btv_range=w["btv_range"], btv_decay=w["btv_decay"], threads=threads)
    for _ in range(warmup):
        fn(m, w["x0"], obs, w["reg_kind"], w["lam"], wts, **kw)
    t0 = time.perf_counter()
    for _ in range(steps):
        fn(m, w["x0"], obs, w["reg_kind"], w["lam"], wts, **kw)
    dt = (time.perf_counter() - t0) / steps
    units = wl.work_units(side, side, w["C"], w["N"])
    desc = ("%s ComputeAllTerms on a %dx%dx%d HR crop, %d frames, %d thread(s), %d evals" %
            ("oracle/_ref" if use_ref else "oracle port", side, side, w["C"], w["N"], threads, steps))
    return units / dt, dt, desc, ("reference-sources+port" if use_ref else "port")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    steps = max(1, min(args.steps, 5))
    warm = max(1, min(args.warmup, 1))
    val, dt, desc, kind = cpu_reference_run(args.config, args.cpu_sample, steps, warm, cores)
    wl = importlib.import_module("super-resolution_b200.workloads")
    cf = wl.CONFIGS[args.config]
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": warm, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": cf["name"], "sample": desc},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc,
                         "note": kind},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def main():
    args = parse_args()
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
    if world > 1:
        # NCCL kernels on a high-priority stream: they must be able to start while the tile kernel
        # of the following gradient slice still fills the SMs
        opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=os.environ.get("SRB_NCCL_PRIO", "1") == "1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank), pg_options=opts)
    assert world == args.gpus or world == 1, "launch with torchrun --nproc-per-node == --gpus"

    cf = wl.CONFIGS[args.config]
    n_local = cf["N"]                       # weak scaling: every rank holds the config's N frames
    n_total = n_local * world
    frames = sharding.frame_shard(n_total, rank, world)
    H, W, C, s = cf["H"], cf["W"], cf["C"], cf["s"]
    shifts_all = wl.default_shifts(n_total, s)
    psf = wl.gaussian_psf(cf["K"], cf["sigma"])

    eng = srb.Engine((n_local, C, H // s, W // s), s, psf, shifts_all[frames], device=local_rank)
    work = wl.make(args.config, forward=lambda k, plane: eng.forward(frames.index(k), plane),
                   N=n_total, frames=frames)
    eng.set_observations(work["lr"])
    eng.set_regularizer(work["reg_kind"], work["lam"], work["btv_range"], work["btv_decay"])
    eng.set_regularizer_rows(*sharding.row_band(H, rank, world))
    eng.set_path({"auto": srb.PATH_AUTO, "reference_order": srb.PATH_REFERENCE_ORDER,
                  "fused": srb.PATH_FUSED}[args.path])
    n = C * H * W
    units = wl.work_units(H, W, C, n_total)
    alg_bytes = wl.algorithmic_bytes(H, W, C, n_local, s, has_reg=True)

    # everything the timed region touches lives on the engine's stream
    stream = torch.cuda.ExternalStream(eng.stream_handle(), device=torch.device("cuda", local_rank))
    x0 = np.ascontiguousarray(work["x0"]).reshape(-1)
    with torch.cuda.stream(stream):
        x_dev = torch.from_numpy(x0).to("cuda", non_blocking=False)
        gc_dev = torch.zeros(n + 1, dtype=torch.float64, device="cuda")
    if world > 1:
        dist.broadcast(x_dev, src=0)        # replicas of the same estimate
    h_x = torch.from_numpy(x0.copy()).pin_memory()
    h_g = torch.empty(n + 1, dtype=torch.float64).pin_memory()
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

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident():
        # one evaluation of the full objective: per-rank partial + (N > 1) the cross-rank sum, either
        # fused into the tile kernel over NVLink peer memory or as a pipelined NCCL allreduce
        if peer is not None:
            peer.evaluate(x_dev)
        else:
            objective.evaluate(x_dev, gc_dev).wait()

    def step_e2e():
        # the call a host solver makes: host x in, host gradient + cost out
        if world == 1:
            cost, _ = eng.eval(h_x.numpy(), out=h_g.numpy()[:n])
            return cost
        if rank == 0:
            x_dev.copy_(h_x, non_blocking=True)
        dist.broadcast(x_dev, src=0)
        if peer is not None:
            peer.evaluate(x_dev)
            if rank == 0:
                h_g.copy_(peer.out[:n + 1], non_blocking=True)
        else:
            objective.evaluate(x_dev, gc_dev).wait()
            if rank == 0:
                h_g.copy_(gc_dev, non_blocking=True)
        stream.synchronize()
        return float(h_g[n]) if rank == 0 else 0.0

    warmup = max(args.warmup, 3)
    with torch.cuda.stream(stream):
        for _ in range(warmup):
            step_resident()
        barrier()
        sampler = ClockSampler(local_rank) if rank == 0 else None
        # ---- device-resident throughput: EXACTLY args.steps steps, barrier + sync on both sides ----
        launches0 = eng.timing()["kernel_launches"]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(args.steps):
            step_resident()
        e1.record(stream)
        barrier()
        ms_total = e0.elapsed_time(e1)
        launches = eng.timing()["kernel_launches"] - launches0

        # ---- the dominant kernel alone: CUDA events recorded by the library around the tile kernel
        #      launch, on the launching stream (srb_set_profiling) ----------------------------------
        eng.set_profiling(True)
        kern_ms = []
        for _ in range(min(max(args.steps, 5), 50)):
            eng.eval_partial_dev(x_dev, gc_dev)
            stream.synchronize()
            kern_ms.append(eng.timing()["last_main_kernel_ms"])
        eng.set_profiling(False)
        kernel_ms = float(np.mean(kern_ms))

        # ---- end to end through the host-facing call ----------------------------------------------
        for _ in range(3):
            step_e2e()
        barrier()
        e_steps = max(3, min(args.steps, 20))
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record(stream)
        for _ in range(e_steps):
            cost = step_e2e()
        f1.record(stream)
        barrier()
        e2e_ms = f0.elapsed_time(f1)
    t = torch.tensor([ms_total, e2e_ms, kernel_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, e2e_ms, kernel_ms = (float(v) for v in t.cpu())
    # keep the same step running until the clock sampler has seen >= ~1 s of load; the number of
    # extra steps is derived from the rank-reduced timings so that every rank runs the same count
    extra = int(max(0.0, 1.2 - (ms_total + e2e_ms) * 1e-3) / max(ms_total / args.steps * 1e-3, 1e-6))
    with torch.cuda.stream(stream):
        for _ in range(min(extra, 20000)):
            step_resident()
        barrier()
    clocks = sampler.stop() if sampler else None
    ms_per_step = ms_total / args.steps
    value = units / (ms_per_step * 1e-3)
    e2e_value = units / (e2e_ms / e_steps * 1e-3)

    if rank == 0:
        peak, peak_src = measured_peaks()
        achieved = alg_bytes / (kernel_ms * 1e-3) / 1e9
        path_name = {1: "reference_order", 2: "fused"}[eng.active_path]
        if path_name == "fused" and eng.zlayout_active:
            path_name = "fused_zlayout"
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": cf["name"], "frames_per_gpu": n_local, "frames_total": n_total,
                       "partition": ("single GPU" if world == 1 else
                                     "frame shard; band-pipelined tile kernel + copy-engine reduce-scatter over "
                                     "NVLink peer memory + sum/all-gather kernel (C*P+1 f64 per step)" if peer is not None else
                                     "frame shard + 1 NCCL allreduce(C*P+1 f64) per step, pipelined in %d slices"
                                     % args.chunks),
                       "kernel_path": path_name,
                       "l2": "inputs larger than L2 (%.0f MB touched per step vs 126 MB L2)" % (alg_bytes / 1e6),
                       "cost_check": cost},
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": e2e_ms / e_steps,
                    "h2d_bytes_per_step": n * 8, "d2h_bytes_per_step": (n + 1) * 8},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                         "kernel_ms": kernel_ms, "algorithmic_bytes": alg_bytes,
                         "kernel": {"fused": "k_tile (fused tile kernel), one launch per evaluation",
                                    "fused_zlayout": "k_tile_z (fused tile kernel, observations in Z layout), "
                                                     "one launch per evaluation"}.get(path_name,
                                                                                      "reference-order kernels")},
        }
        prof = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(prof):
            try:
                line["roofline"]["traffic"] = json.load(open(prof)).get(path_name)
            except Exception:
                pass
        if args.solve_iters > 0 and world == 1:
            # the whole inner solve behind one C-ABI call; x crosses PCIe once each way
            import time
            t0 = time.perf_counter()
            _, rep = eng.cg_minimize(h_x.numpy(), maxits=args.solve_iters)
            dt = time.perf_counter() - t0
            line["solve"] = {"api": "srb_cg_minimize", "iterations": rep["iterations"],
                             "evaluations": rep["num_evaluations"], "termination_type": rep["termination_type"],
                             "seconds": dt, "value": units * rep["num_evaluations"] / dt, "unit": UNIT,
                             "h2d_bytes": n * 8, "d2h_bytes": n * 8, "final_cost": rep["final_cost"]}
        if not args.no_cpu_baseline and world == 1:
            cores = os.cpu_count() or 1
            val, dt, desc, kind = cpu_reference_run(args.config, args.cpu_sample, 3, 1, cores)
            line["cpu_baseline"] = {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": desc, "note": kind}
        print(json.dumps(line))
    # release every tensor that lives on the engine's stream before the stream goes away
    if peer is not None:
        peer.close()
    del objective, peer, x_dev, gc_dev, h_x, h_g, t
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    eng.close()


if __name__ == "__main__":
    main()
