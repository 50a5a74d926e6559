#!/usr/bin/env python
"""bench.py - headline benchmark of the Solaris force + integrator hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--n BODIES]

Workload (BASELINE.json metric, SURVEY.md §8d config "H"): a synthetic self-gravitating disk of
N = 10^6 bodies (star + protoplanets, astrocentric frame, no nebula), integrated with RKF7(8).
One "step" = one RungeKuttaFehlberg78::Driver call on the device-resident system = 13 force
evaluations (more if an attempt is rejected) + the stage / error kernels.

metric  = fp64 pair interactions per second (ordered sink-source pairs, whole job, all GPUs)
value   = pairs evaluated in the K timed steps / device time (CUDA events, max over ranks)
e2e     = the same through the host-buffer path: every step uploads y0 from pinned host memory,
          runs the Driver, downloads the new y0 (what the drop-in Driver does for the reference's
          Simulator, which owns host arrays)
roofline= the pair kernel against the FP64 FMA peak measured in this run (MEASURED_PEAKS.json has no
          fp64 entry): achieved = 20 flop/pair x pairs per launch / mean launch duration
cpu_baseline = the plain-C oracle port on all host cores, sink subset of the SAME system
--impl reference = the UNMODIFIED compiled reference (oracle/_ref) Acceleration::Compute, one replica
          per host thread, bounded sample N_cpu = 16384.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "fp64_pair_interactions_per_s"
UNIT = "pairs/s"
FLOP_PER_PAIR = 20.0


def workload_name(n):
    return (f"H: synthetic self-gravitating disk, N={n} (1 star + {n - 1} protoplanets, m~U(0.001,0.1) M_earth, "
            f"a~U(5,6) AU, e~U(0,0.1)), astrocentric, no nebula, RKF78, nearest-neighbour side outputs produced by the last "
            f"stage of every step (the only ones Simulator::CheckEvent can observe)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu), "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thr = threading.Thread(target=self._pump, daemon=True)
            self.thr.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------------------
# reference arm: the compiled, unmodified reference on the host cores
# --------------------------------------------------------------------------------------------------
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from solaris_b200 import synth
    import oraclelib
    n_cpu = 16384
    cores = os.cpu_count() or 1
    line = {"impl": "reference", "metric": METRIC, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": {"workload": workload_name(args.n), "timing": "host steady_clock"},
            "gpu_launches": 0}
    sysm = synth.massive_disk(n_cpu)
    pairs = synth.pairs_per_eval(sysm.counts, False)
    if oraclelib.reference_available():
        kind = "reference"
        replicas = [oraclelib.Reference(sysm, False, None) for _ in range(cores)]

        def one(rep):
            rep.time_compute(0.0, 1)
    else:
        kind = "port"
        replicas = [oraclelib.Oracle(sysm, False, None) for _ in range(cores)]

        def one(rep):
            rep.compute(0.0, sysm.y0, 0)

    def step():
        th = [threading.Thread(target=one, args=(r,)) for r in replicas]
        for t in th:
            t.start()
        for t in th:
            t.join()

    for _ in range(max(1, min(args.warmup, 1))):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    value = pairs * cores * args.steps / dt
    sample = (f"{cores} independent replicas (one per host thread) of Acceleration::Compute on a self-gravitating disk of "
              f"N_cpu={n_cpu} bodies ({pairs:.3e} pairs per evaluation), astrocentric; the reference itself is single-threaded")
    line.update({"value": value, "ms_per_step": 1e3 * dt / args.steps,
                 "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
                 "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
    print(json.dumps(line), flush=True)
    return 0


# --------------------------------------------------------------------------------------------------
# CPU baseline leg of the main arm (rank 0, N=1 only)
# --------------------------------------------------------------------------------------------------
def cpu_baseline(sysm, pairs_per_sink):
    import oraclelib
    cores = os.cpu_count() or 1
    o = oraclelib.Oracle(sysm, False, None)
    # calibrate on a few rows, then size the sample for ~12 s of wall time on all cores
    t0 = o.time_gravity_rows(1, 2, 1, 1)                       # fixed cost: the serial rm3 pass over all bodies
    cal = 64 * cores
    t = max(o.time_gravity_rows(1, 1 + cal, cores, 1) - t0, 1e-6)
    rows = int(max(cal, min(sysm.n - 1, (15.0 / t) * cal)))
    t = o.time_gravity_rows(1, 1 + rows, cores, 1)
    value = rows * pairs_per_sink / t
    out = {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
           "sample": f"oracle/oracle.c row-subset restatement of GravityAC (Acceleration.cpp:268-326): first {rows} sinks against "
                     f"all {sysm.n - 1} sources of the benchmark system, {cores} pthreads, {t:.1f} s"}
    if oraclelib.reference_available():
        from solaris_b200 import synth
        small = synth.massive_disk(16384)
        r = oraclelib.Reference(small, False, None)
        tt = r.time_compute(0.0, 3)
        out["reference_1core"] = {"value": synth.pairs_per_eval(small.counts, False) / tt, "unit": UNIT, "cores": 1,
                                  "kind": "reference", "sample": "compiled reference Acceleration::Compute, N_cpu=16384, median of 3"}
    return out


# --------------------------------------------------------------------------------------------------
# multi-GPU correctness inside the driver-run record (WORLD_SIZE > 1, before the timed region)
# --------------------------------------------------------------------------------------------------
def multi_gpu_check(ctx, rank, world, dist):
    """20 000 self-gravitating bodies sharded over all ranks through the symmetric kernel's multi-GPU path (rounds dealt to
    the ranks, partial sums combined over NVLink): every rank compares >= 1024/world of its own sinks with the
    extended-precision row oracle (north star: 1e-13) and their nearest neighbours with the reference's row arithmetic.
    The oracle is used as the CHECKER only, outside every timed region."""
    import numpy as np
    import torch
    from solaris_b200 import synth
    import oraclelib
    n = 20000
    s = synth.massive_disk(n)
    ctx.set_frame(False); ctx.set_nn_tracking(2); ctx.set_pair_algorithm(2); ctx.set_bodies(s); ctx.set_nebula(None)
    a = np.zeros((n, 6))
    ctx.compute(0.0, s.y0, 0, out=a)                      # collective; fills this rank's rows
    nn = ctx.download(capi_mod().NN_INDEX)
    lo, hi = ctx.shard_range()
    lo = max(lo, 1)
    rows = np.unique(np.random.default_rng(100 + rank).integers(lo, hi, max(1024 // world, 128))).astype(np.int32) if hi > lo else np.zeros(0, np.int32)
    o = oraclelib.Oracle(s, False, None)
    err, bad_nn = 0.0, 0
    if len(rows):
        ex = o.gravity_rows_exact(s.y0, rows, threads=max(1, (os.cpu_count() or 1) // world))
        err = float((np.abs(a[rows, 3:] - ex).max(axis=1) / np.sqrt((ex ** 2).sum(axis=1))).max())
        for i in rows[:64]:
            o.gravity_rows(s.y0, int(i), int(i) + 1, 1)
            bad_nn += int(nn[i] != o.side()[1][i])
    t = torch.tensor([err, float(bad_nn), float(len(rows))], dtype=torch.float64)
    tm = t.clone(); dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    ts = t.clone(); dist.all_reduce(ts, op=dist.ReduceOp.SUM)
    return {"ranks": world, "bodies": n, "rows": int(ts[2].item()), "max_err": float(tm[0].item()), "tolerance": 1.0e-13,
            "nn_mismatches": int(ts[1].item()), "path": "sol_compute on the sharded context (symmetric kernel, rounds dealt to the ranks)",
            "checker": "oracle_gravity_rows_exact (long double, compensated) + reference row arithmetic for indexOfNN",
            "ok": bool(tm[0].item() <= 1.0e-13 and ts[1].item() == 0 and ts[2].item() >= 1024 * 0.9)}


def capi_mod():
    from solaris_b200 import capi
    return capi


# --------------------------------------------------------------------------------------------------
# BASELINE.json configs C1..C5 beside the headline (rank 0 reports; C4 / C5 run on all ranks)
# --------------------------------------------------------------------------------------------------
INTEG_NAME = {0: "DormandPrince", 1: "RungeKutta4", 3: "RungeKuttaFehlberg78"}
# algorithmic bytes per body of the stage + solution / error kernels
ALG_BYTES_ATTEMPT = {3: 4464.0, 1: 720.0, 0: 1800.0}


def _h0(s):
    import numpy as np
    from solaris_b200 import synth
    r = np.sqrt((s.y0[1:, :3] ** 2).sum(axis=1)); v2 = (s.y0[1:, 3:] ** 2).sum(axis=1)
    mu = synth.GAUSS2 * (1.0 + s.mass[1:])
    a = 1.0 / (2.0 / r - v2 / mu)
    return float((2 * np.pi * np.sqrt(a ** 3 / mu)).min() / 50000.0)     # Simulator::MainIntegration, Simulator.cpp:435


class stdout_to_stderr:
    """The compiled reference announces the drag regime on stdout (Acceleration.cpp:362-365) and NCCL its version; stdout
    belongs to the ONE JSON line."""
    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)

    def __exit__(self, *exc):
        os.dup2(self.saved, 1)
        os.close(self.saved)


def reference_steps_per_s(s, integ, neb, h0, budget_s, max_steps):
    with stdout_to_stderr():
        return _reference_steps_per_s(s, integ, neb, h0, budget_s, max_steps)


def _reference_steps_per_s(s, integ, neb, h0, budget_s, max_steps):
    """The compiled reference's own Driver on ONE host core (the reference is single-threaded), same system, same h0."""
    import oraclelib
    if not oraclelib.reference_available():
        return None
    r = oraclelib.Reference(s, False, neb, integ)
    t, h = 0.0, h0
    _, t, h, *_ = r.step(integ, t, h)                     # first call allocates
    n, t0 = 0, time.perf_counter()
    while n < max_steps and (n == 0 or time.perf_counter() - t0 < budget_s):
        _, t, h, *_ = r.step(integ, t, h)
        n += 1
    dt = time.perf_counter() - t0
    return {"steps_per_s": n / dt, "steps_timed": n, "cores": 1, "kind": "reference",
            "sample": f"{n} Driver calls of the compiled reference on this system, one core (the reference is single-threaded)"}


def run_configs(ctx, world, rank, barrier, hbm_peak, fp64_peak, ref_pairs_per_s_1core, only=None):
    """steps/s, pairs/s and the family roofline of BASELINE.json's five configs on the device-resident path, each with the
    compiled reference's steps/s on one host core of the same box beside it (extrapolated where a reference step would take
    hours, and labelled so).  N > 1: only the configs that shard (C4, C5)."""
    import numpy as np
    from solaris_b200 import capi, synth
    import oraclelib
    out = {}
    neb = oraclelib.default_nebula()
    specs = [
        ("C1", "TestCases/SunJupiter: Sun + Jupiter, RKF78", synth.sun_jupiter, capi.RUNGE_KUTTA_FEHLBERG78, None, 20000, False),
        ("C2", "TestCases/SolarSystem: Sun + 8 planets, RKF78", synth.solar_system, capi.RUNGE_KUTTA_FEHLBERG78, None, 20000, False),
        ("C3", "Sun + Jupiter + 10^5 planetesimals with gas drag, RK4", lambda: synth.planetesimal_drag(100_000), capi.RUNGE_KUTTA4, neb, 300, False),
        ("C4", "Sun + Jupiter + Saturn + 10^6 Trojan test particles, DormandPrince RKN7(6)", lambda: synth.trojans(1_000_000), capi.DORMAND_PRINCE, None, 100, True),
        ("C5", "1 star + 262143 protoplanets with type-I migration, default nebula, RKF78", lambda: synth.massive_disk(262_144, migration=True), capi.RUNGE_KUTTA_FEHLBERG78, neb, 2, True),
    ]
    for key, desc, make, integ, nebula, nsteps, shards in specs:
        if only and key.lower() not in only:
            continue
        if world > 1 and not shards:
            continue
        s = make()
        pairs_eval = synth.pairs_per_eval(s.counts, False)
        ctx.set_frame(False); ctx.set_nn_tracking(2); ctx.set_pair_algorithm(2); ctx.set_bodies(s); ctx.set_nebula(nebula)
        h0 = _h0(s)
        # warm-up, then nsteps in ONE sol_run call (persistent kernel for C1 / C2, host loop inside the library otherwise)
        rc, a, _ = ctx.run(integ, 0.0, h0, max(2, min(nsteps // 10, 200)))
        if rc != 0:
            raise SystemExit(f"{key}: driver failed: " + ctx.last_error())
        ctx.profile_read(reset=True); ctx.profile_enable(True)
        l0 = ctx.launch_count()
        barrier()
        t0 = time.perf_counter()
        rc, a, _ = ctx.run(integ, a.time, a.h_next, nsteps)
        barrier()
        dt = time.perf_counter() - t0
        if rc != 0:
            raise SystemExit(f"{key}: driver failed: " + ctx.last_error())
        ms, cnt = ctx.profile_read(reset=True)
        ctx.profile_enable(False)
        ne = {3: 13, 1: 4, 0: 9}[integ]
        evals = a.attempts * (ne - 1) + a.steps                  # k0 once per step, the other stages once per attempt
        rec = {"workload": desc, "bodies": int(s.n), "integrator": INTEG_NAME[integ], "n_gpus": world, "steps": int(a.steps),
               "attempts": int(a.attempts), "steps_per_s": a.steps / dt, "us_per_step": 1e6 * dt / max(a.steps, 1),
               "pairs_per_eval": pairs_eval, "pairs_per_s": pairs_eval * evals / dt,
               "gpu_launches": ctx.launch_count() - l0,
               "timing": "host wall clock around one sol_run call (barrier + stream synchronise on both sides)", "h0_days": h0}
        if key in ("C3", "C4"):
            # HBM family: the attempt kernels of the tracer path against the reference's formulation of the same work
            n_rank = s.n / world
            alg = ALG_BYTES_ATTEMPT[integ] * a.attempts * n_rank + (112.0 * evals * s.counts[5] / world if key == "C3" else 0.0)
            fam_ms = ms[5] + ms[2] + ms[3] + ms[4]
            gbs = alg / (fam_ms * 1e-3) / 1e9 if fam_ms > 0 else None
            peak = hbm_peak if hbm_peak else 6650.0
            rec["roofline"] = {"bound": "hbm", "kernel": "tracer_attempt_kernel (+ the one-warp kernel for the massive bodies)",
                               "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak if gbs else None,
                               "algorithmic_bytes": alg, "kernel_ms": fam_ms,
                               "note": ("algorithmic bytes are SURVEY.md 8(d)'s count of the reference's array passes (720 N per RK4 step, "
                                        "~1800 N per RKN attempt, 112 B of drag operands per planetesimal and evaluation); the fused kernel "
                                        "keeps the k-vectors in registers and moves ~100 B per body and attempt, so frac can exceed 1: the "
                                        "kernel is bound by FP64 issue and latency, not by HBM (profiles/)")}
        elif key == "C5":
            pair_ms = ms[0]
            pairs_rank = pairs_eval * evals / world
            ach = FLOP_PER_PAIR * pairs_rank / (pair_ms * 1e-3) / 1e12 if pair_ms > 0 else None
            rec["roofline"] = {"bound": "fp64", "kernel": "sol::sym_pair_kernel", "achieved": ach, "peak": fp64_peak, "unit": "TFLOP/s",
                               "frac": ach / fp64_peak if ach and fp64_peak else None, "share_of_step": pair_ms * 1e-3 / dt,
                               "peak_source": "in-run DFMA probe (sol_measure_fp64_peak)"}
            stage_ms = ms[2] + ms[3] + ms[4]
            alg = (4464.0 * a.attempts + 144.0 * a.steps + 120.0 * evals) * s.n / world
            rec["roofline_hbm"] = {"bound": "hbm", "kernels": "finalize (+ type-I migration, + next stage) + stage + solution/error",
                                   "achieved": alg / (stage_ms * 1e-3) / 1e9 if stage_ms > 0 else None, "peak": hbm_peak,
                                   "unit": "GB/s", "algorithmic_bytes": alg, "kernel_ms": stage_ms}
        else:
            rec["roofline"] = {"bound": "latency", "kernel": "sol::cp_run_kernel (one persistent warp for the whole call)",
                               "note": "a 2- or 9-body step is one dependency chain; neither FP64 throughput nor HBM bandwidth is in play"}
        if rank == 0:
            if key in ("C1", "C2"):
                rec["reference_cpu"] = reference_steps_per_s(s, integ, nebula, h0, 2.0, 20000)
            elif key in ("C3", "C4"):
                rec["reference_cpu"] = reference_steps_per_s(s, integ, nebula, h0, 4.0, 3)
            elif ref_pairs_per_s_1core:
                per_step = pairs_eval * evals / max(a.steps, 1)
                rec["reference_cpu"] = {"steps_per_s": ref_pairs_per_s_1core / per_step, "cores": 1, "kind": "reference",
                                        "extrapolated": True,
                                        "sample": "EXTRAPOLATED: the compiled reference's measured pairs/s on one core (N_cpu = 16384) divided by "
                                                  f"the {per_step:.3e} pair interactions of one step of this system (a real step would take hours)"}
            if rec.get("reference_cpu"):
                rec["speedup_vs_reference_1core"] = rec["steps_per_s"] / rec["reference_cpu"]["steps_per_s"]
        barrier()
        out[key] = rec
    return out


# --------------------------------------------------------------------------------------------------
# main arm
# --------------------------------------------------------------------------------------------------
def run_b200(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from solaris_b200 import capi, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - solaris_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # torch.distributed is only the rendezvous (unique-id broadcast, barriers, max over ranks) -> gloo;
        # the data path's collectives are the library's own NCCL communicator (sol_dist_init)
        dist.init_process_group(backend="gloo", rank=rank, world_size=world)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    n = args.n
    sysm = synth.massive_disk(n)
    pairs_eval = synth.pairs_per_eval(sysm.counts, False)

    ctx = capi.Context(local_rank)
    stream = torch.cuda.Stream()
    ctx.set_stream(stream.cuda_stream)
    if world > 1:
        uid = [capi.Context.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        # NCCL announces its version on stdout when the communicator is created; keep stdout for the ONE JSON line
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            ctx.dist_init(rank, world, uid[0])
        finally:
            os.dup2(saved, 1)
            os.close(saved)
    mgc = multi_gpu_check(ctx, rank, world, dist) if world > 1 else None
    ctx.set_frame(False)
    # nn mode 2: indexOfNN / distanceOfNN are produced by the LAST stage of each step - exactly the values
    # the reference leaves behind for CheckEvent; earlier stages' NN arrays are
    # dead stores in the reference (overwritten before anything can read them).
    ctx.set_nn_tracking(args.nn_mode)
    ctx.set_pair_algorithm(0 if args.ordered else 1)
    ctx.set_bodies(sysm)
    ctx.set_nebula(None)

    fp64_peak = ctx.measure_fp64_peak()        # TFLOP/s, this GPU, this run

    # Simulator::MainIntegration: h0 = ShortestPeriod()/50000 (Simulator.cpp:435)
    r = np.sqrt((sysm.y0[1:, :3] ** 2).sum(axis=1))
    v2 = (sysm.y0[1:, 3:] ** 2).sum(axis=1)
    mu = synth.GAUSS2 * (1.0 + sysm.mass[1:])
    a = 1.0 / (2.0 / r - v2 / mu)
    period = 2.0 * np.pi * np.sqrt(a ** 3 / mu)
    h0 = float(period.min() / 50000.0)

    INT = capi.RUNGE_KUTTA_FEHLBERG78
    t, h = 0.0, h0
    for _ in range(args.warmup):
        rc, t, h, hd, att, em, ev, pr = ctx.step(INT, t, h)
        if rc != 0:
            raise SystemExit("driver failed: " + ctx.last_error())

    # ---- timed region: K steps on the device-resident system ----
    sampler = ClockSampler(local_rank)
    ctx.profile_read(reset=True)
    ctx.profile_enable(True)
    launches0 = ctx.launch_count()
    barrier()
    sampler.start()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record(stream)
        pairs_total = 0.0; evals_total = 0.0; attempts_total = 0
        for _ in range(args.steps):
            rc, t, h, hd, att, em, ev, pr = ctx.step(INT, t, h)
            if rc != 0:
                raise SystemExit("driver failed: " + ctx.last_error())
            pairs_total += pr; evals_total += ev; attempts_total += att
        e1.record(stream)
    barrier()
    clocks = sampler.stop()
    ms = e0.elapsed_time(e1)
    launches = ctx.launch_count() - launches0
    prof_ms, prof_n = ctx.profile_read(reset=True)
    ctx.profile_enable(False)
    if world > 1:
        tms = torch.tensor([ms], dtype=torch.float64)
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        ms = float(tms.item())
    value = pairs_total / (ms * 1e-3)          # pairs_total counts the WHOLE system (all ranks' sinks)

    # ---- e2e: host buffers in and out of every step ----
    y_host = torch.empty((n, 6), dtype=torch.float64).pin_memory()
    e2e_steps = max(1, min(args.steps, 2))
    barrier()
    if world > 1:
        ctx.gather_state()
    y_host.copy_(torch.from_numpy(ctx.download(capi.Y0)))
    barrier()
    e2 = torch.cuda.Event(enable_timing=True); e3 = torch.cuda.Event(enable_timing=True)
    e2e_pairs = 0.0
    with torch.cuda.stream(stream):
        e2.record(stream)
        for _ in range(e2e_steps):
            ctx.lib.sol_upload(ctx.h, capi.Y0, y_host.data_ptr())
            rc, t, h, hd, att, em, ev, pr = ctx.step(INT, t, h)
            if rc != 0:
                raise SystemExit("driver failed: " + ctx.last_error())
            if world > 1:
                ctx.gather_state()
            ctx.lib.sol_download(ctx.h, capi.Y0, y_host.data_ptr())
            e2e_pairs += pr
        e3.record(stream)
    barrier()
    ms_e2e = e2.elapsed_time(e3)
    if world > 1:
        tms = torch.tensor([ms_e2e], dtype=torch.float64)
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        ms_e2e = float(tms.item())
    e2e_value = e2e_pairs / (ms_e2e * 1e-3)

    shard_h = ctx.shard_range()       # of the headline system (the configs block below loads other systems)
    # ---- side leg: the same step with nearest-neighbour outputs in EVERY evaluation (the reference's habit) ----
    nn_all = None
    if args.nn_mode != 1 and not args.no_nn_leg:
        ctx.set_nn_tracking(1)
        barrier()
        e4 = torch.cuda.Event(enable_timing=True); e5 = torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            e4.record(stream)
            rc, t, h, hd, att, em, ev, pr = ctx.step(INT, t, h)
            e5.record(stream)
        barrier()
        if rc != 0:
            raise SystemExit("driver failed: " + ctx.last_error())
        ms_nn = e4.elapsed_time(e5)
        if world > 1:
            tms = torch.tensor([ms_nn], dtype=torch.float64)
            dist.all_reduce(tms, op=dist.ReduceOp.MAX)
            ms_nn = float(tms.item())
        nn_all = {"value": pr / (ms_nn * 1e-3), "unit": UNIT, "steps": 1, "ms_per_step": ms_nn, "force_evals": ev,
                  "note": "indexOfNN / distanceOfNN produced by all 13 evaluations of the attempt instead of the last one only"}
        ctx.set_nn_tracking(args.nn_mode)

    # ---- BASELINE.json's other configs (all ranks take part in the ones that shard) ----
    configs = None
    if not args.no_configs:
        ref_1core = None
        if rank == 0:
            import oraclelib
            if oraclelib.reference_available():
                small = synth.massive_disk(16384)
                ref_1core = synth.pairs_per_eval(small.counts, False) / oraclelib.Reference(small, False, None).time_compute(0.0, 2)
        lo_h, hi_h = ctx.shard_range()
        hbm_pk = None
        try:
            hbm_pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs")
        except Exception:
            pass
        only = [c.strip().lower() for c in args.configs.split(",")] if args.configs else None
        configs = run_configs(ctx, world, rank, barrier, hbm_pk, fp64_peak, ref_1core, only)
        if ref_1core and configs is not None:
            configs["reference_pairs_per_s_1core_n16384"] = ref_1core

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel (pair kernel), live numbers of the timed region ----
    # achieved = 20 flop x ordered pairs credited to this rank's pair-kernel launches / their summed duration.
    # With the symmetric kernel one launch covers up to 32 rounds of block pairs; every unordered pair is
    # evaluated once (20 FP64 instructions) and credited as the two ordered pairs the reference evaluates.
    lo, hi = shard_h
    sym = (not args.ordered) and (n - 1) >= 4096
    pairs_rank = pairs_total / world if sym else pairs_total * float(max(hi, 1) - max(lo, 1)) / float(n - 1)
    pair_ms_total = prof_ms[0]
    achieved = FLOP_PER_PAIR * pairs_rank / (pair_ms_total * 1e-3) / 1e12 if pair_ms_total > 0 else None
    instr_per_pair = 10 if sym else 16
    roofline = {"bound": "fp64", "kernel": "sol::sym_pair_kernel" if sym else "sol::pair_kernel", "achieved": achieved,
                "peak": fp64_peak, "unit": "TFLOP/s", "frac": (achieved / fp64_peak) if achieved else None, "traffic": None,
                "peak_source": "in-run dependent-free DFMA probe on all SMs (sol_measure_fp64_peak); MEASURED_PEAKS.json carries no fp64 figure",
                "flop_per_pair": FLOP_PER_PAIR, "pairs_per_launch": pairs_rank / max(prof_n[0], 1),
                "ms_per_launch": pair_ms_total / max(prof_n[0], 1), "launches_timed": prof_n[0],
                "ms_per_force_eval": pair_ms_total / max(evals_total, 1),
                "share_of_step": pair_ms_total / ms if ms > 0 else None,
                "fp64_instr_per_pair": instr_per_pair,
                # the like-for-like figure: nearest-neighbour outputs in all 13 evaluations, as the reference's loop literally
                # does (20 flop per pair INCLUDE the compare + select); `frac` has them in the last stage only
                "frac_nn_every_eval": (FLOP_PER_PAIR * nn_all["value"] / 1e12 / fp64_peak) if nn_all else None,
                "pipe_frac": (achieved / fp64_peak) * instr_per_pair * 2 / FLOP_PER_PAIR if achieved else None,
                "note": ("symmetric kernel: each unordered pair is evaluated once with 20 FP64 instructions and credited as 2 ordered "
                         "pairs x 20 flop (the reference's count), so frac can exceed the 62.5 % ceiling of the ordered kernel"
                         if sym else "ordered kernel: 16 FP64 instructions per ordered pair")}
    traffic_file = os.path.join(ROOT, "profiles", "pair_kernel_traffic.json")
    if os.path.exists(traffic_file):
        try:
            roofline["traffic"] = json.load(open(traffic_file)).get(f"n{n}", {}).get("dram_bytes_per_launch")
        except Exception:
            pass

    # HBM-bound kernel families of the same timed region.  The stage combination of stage s+1 is formed by the finalize
    # kernel of evaluation s (one launch less per stage), so the three families are reported together:
    # finalize (+ fused stage) + the stand-alone stage / yscale launches + solution/error norm.
    # Algorithmic bytes of SURVEY.md §8(d): 4464 N per RKF78 attempt + 144 N once per step for yscale, plus the finalize's
    # own 6 state + 3 pair-sum + 6 derivative doubles = 120 N per evaluation.
    hbm_peak = None
    try:
        hbm_peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs")
    except Exception:
        pass
    n_rank = float(hi - lo)
    stage_ms = prof_ms[2] + prof_ms[3] + prof_ms[4]
    stage_bytes = (4464.0 * attempts_total + 144.0 * args.steps + 120.0 * evals_total) * n_rank
    stage_gbs = stage_bytes / (stage_ms * 1e-3) / 1e9 if stage_ms > 0 else None
    roofline_hbm = {"bound": "hbm", "kernels": "finalize_kernel (incl. the next stage's combination) + rk_stage_kernel<NT> + yscale_kernel + rkf78_final_kernel", "achieved": stage_gbs,
                    "peak": hbm_peak if hbm_peak else 6650.0, "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if hbm_peak else "fallback 6.65 TB/s (of fallback)",
                    "unit": "GB/s", "frac": (stage_gbs / (hbm_peak if hbm_peak else 6650.0)) if stage_gbs else None,
                    "algorithmic_bytes": stage_bytes, "ms": stage_ms, "share_of_step": stage_ms / ms if ms > 0 else None}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(n), "integrator": "RungeKuttaFehlberg78", "bodies": n, "nn_mode": args.nn_mode,
                   "pair_algorithm": "ordered" if args.ordered else "symmetric (unordered pairs once)",
                   "parallelism": f"sinks sharded over {world} GPU(s), sources replicated" if world > 1 else "single GPU",
                   "l2": "inputs larger than L2 (13 k-arrays x 48 MB + partial sums)", "h0_days": h0},
        "steps_per_s": args.steps / (ms * 1e-3), "force_evals": evals_total, "attempts": attempts_total,
        "pairs_per_eval": pairs_eval,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 48 * n, "d2h_bytes_per_step": 48 * n,
                "steps": e2e_steps, "ms_per_step": ms_e2e / e2e_steps,
                "path": "sol_upload(Y0, pinned host) -> sol_step(RKF78) -> sol_download(Y0, pinned host)"},
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": roofline,
        "roofline_hbm": roofline_hbm,
        "nn_every_evaluation": nn_all,
        "configs": configs,
        "multi_gpu_check": mgc,
        "kernel_ms": {"pair": prof_ms[0], "source_prep_indirect": prof_ms[1], "finalize": prof_ms[2], "rk_stage": prof_ms[3],
                      "solution_error": prof_ms[4], "misc": prof_ms[5]},
    }
    if world > 1:
        # family 5 on a sharded context = the NCCL collectives of the timed steps on rank 0 (all-gather of the staged source
        # slices + reduce-scatter of the partial sums per evaluation); their device time includes the wait for the slowest rank
        evals = max(evals_total, 1.0)
        recv_mb = (world - 1) / world * (32.0 * n + 24.0 * n) / 1e6
        line["collectives"] = {
            "ms_total": prof_ms[5], "ms_per_evaluation": prof_ms[5] / evals, "share_of_step": prof_ms[5] / ms,
            "nvlink_mbytes_received_per_rank_and_evaluation": recv_mb,
            "note": "all-gather: every rank receives the other ranks' {x,y,z,m} slices (32 B per body); reduce-scatter of three "
                    "partial-sum planes (ring: (R-1)/R x 24 B per body sent and received per rank); byte counts are algorithmic",
            "uninstrumented_share_of_step": max(0.0, 1.0 - sum(prof_ms) / ms)}
    if world == 1 and not args.no_cpu:
        line["cpu_baseline"] = cpu_baseline(sysm, float(n - 2))
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--bodies", "--n", dest="n", type=int, default=1_000_000)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--nn-mode", type=int, default=2, help="1: NN arrays in every evaluation, 2: last stage only, 0: never")
    ap.add_argument("--no-nn-leg", action="store_true", help="skip the extra step with NN outputs in every evaluation")
    ap.add_argument("--ordered", action="store_true", help="force the ordered pair kernel (one evaluation per ordered pair)")
    ap.add_argument("--no-configs", action="store_true", help="skip the block with BASELINE.json's configs C1..C5")
    ap.add_argument("--configs", default="", help="comma-separated subset of c1,c2,c3,c4,c5 for that block")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
