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

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel (pair kernel), live numbers of the timed region ----
    # achieved = 20 flop x ordered pairs credited to this rank's pair-kernel launches / their summed duration.
    # With the symmetric kernel one launch covers up to 32 rounds of block pairs; every unordered pair is
    # evaluated once (20 FP64 instructions) and credited as the two ordered pairs the reference evaluates.
    lo, hi = ctx.shard_range()
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
        "kernel_ms": {"pair": prof_ms[0], "source_prep_indirect": prof_ms[1], "finalize": prof_ms[2], "rk_stage": prof_ms[3],
                      "solution_error": prof_ms[4], "misc": prof_ms[5]},
    }
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
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
