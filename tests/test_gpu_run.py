"""sol_run: many Driver steps per call (include/solaris_b200.h).  The loop it replaces is Simulator::Integrate's
(Solaris/Simulator.cpp:131-170) with Simulator::DecisionMaking (:181-248) between the steps; `py_run` below restates that
loop on top of the single-step entry points (sol_step, sol_detect_events, sol_flush_tiny), which are themselves
parity-tested against the oracle.  Systems of <= 32 massive bodies run sol_run as ONE persistent kernel launch."""
import numpy as np
import pytest

from solaris_b200 import capi, synth
from helpers import configure, orbital_elements_ae, total_energy
from oraclelib import Oracle

pytestmark = pytest.mark.gpu


def py_run(ctx, integrator, time, h_next, max_steps, length=1e300, output=1e300, last_save=0.0, millenium_days=0.0,
           ejection=0.0, hit_centrum=0.0, collision_factor=0.0, step_counter=0, flush_every=100):
    steps, reason, attempts, recs, ev = 0, capi.RUN_MAX_STEPS, 0, [], [0, 0, 0]
    h_did = 0.0
    while steps < max_steps:
        h_trial = h_next
        rc, time, h_next, h_did, att, em, *_ = ctx.step(integrator, time, h_next)
        assert rc == 0
        attempts += att
        steps += 1
        step_counter += 1
        recs.append((time, h_did, h_next, h_trial))
        if ejection > 0 or hit_centrum > 0 or collision_factor > 0:
            lists = ctx.detect_events(ejection, hit_centrum, collision_factor)
            ev = [len(x) for x in lists]
            if sum(ev) > 0:
                reason = capi.RUN_EVENT
                break
        ls = last_save + h_did
        actual = millenium_days + time
        if abs(actual) >= abs(length):
            reason = capi.RUN_END
            break
        hn = h_next
        if abs(actual + hn) > abs(length):
            hn = length - actual
        if abs(ls) >= abs(output):
            reason = capi.RUN_SAVE
            break
        if abs(ls + hn) > abs(output):
            hn = output - ls
        last_save, h_next = ls, hn
        if flush_every > 0 and step_counter % flush_every == 0:
            ctx.flush_tiny(1.0e-50)
    return dict(time=time, h_next=h_next, h_did=h_did, last_save=last_save, steps=steps, reason=reason, attempts=attempts,
                step_counter=step_counter, ev=ev, recs=np.array(recs))


SMALL = [("sun-jupiter", lambda: synth.mixed([1, 1, 0, 0, 0, 0, 0], migration=False)),
         ("solar-system", lambda: synth.solar_system())]


@pytest.mark.parametrize("kernel", [1, 3], ids=["component-parallel", "body-per-lane"])
@pytest.mark.parametrize("case", SMALL, ids=[c[0] for c in SMALL])
def test_run_rk4_is_bit_identical_to_the_step_loop(ctx, case, kernel):
    """RK4 has no step-size formula, so the persistent kernel must reproduce the step loop bit for bit: state, previous
    state, times, the clamps of DecisionMaking (output = 2.05 steps forces a clamped step), the flush, the stop reason."""
    s = case[1]()
    h = 0.37
    kw = dict(length=1.0e9, output=2.05 * h * 60, flush_every=7)
    configure(ctx, s, False, None)
    ref = py_run(ctx, capi.RUNGE_KUTTA4, 0.0, h, 300, **kw)
    y_ref, yp_ref = ctx.download(capi.Y0), ctx.download(capi.Y)
    side_ref = (ctx.download(capi.RM3), ctx.download(capi.NN_INDEX), ctx.download(capi.NN_DISTANCE))
    configure(ctx, s, False, None)
    ctx.set_small_system_kernel(kernel)
    n0 = ctx.launch_count()
    rc, a, rec = ctx.run(capi.RUNGE_KUTTA4, 0.0, h, 300, records=True, **kw)
    ctx.set_small_system_kernel(1)
    assert rc == 0 and ctx.launch_count() - n0 == 1, "one persistent launch"
    assert a.stop_reason == ref["reason"] == capi.RUN_SAVE and a.steps == ref["steps"] > 100
    assert (a.time, a.h_next, a.h_did, a.last_save, a.step_counter) == (ref["time"], ref["h_next"], ref["h_did"], ref["last_save"], ref["step_counter"])
    assert np.array_equal(rec, ref["recs"])
    assert np.array_equal(ctx.download(capi.Y0), y_ref) and np.array_equal(ctx.download(capi.Y), yp_ref)
    for got, want in zip((ctx.download(capi.RM3), ctx.download(capi.NN_INDEX), ctx.download(capi.NN_DISTANCE)), side_ref):
        assert np.array_equal(got, want)


@pytest.mark.parametrize("case", SMALL, ids=[c[0] for c in SMALL])
@pytest.mark.parametrize("integrator", [capi.RUNGE_KUTTA_FEHLBERG78, capi.DORMAND_PRINCE])
def test_run_adaptive_against_oracle(ctx, case, integrator):
    """north_star horizon check through sol_run: C1 400 / C2 200 accepted adaptive steps in ONE launch against the
    free-running oracle: identical accepted / rejected counts, energy and (a, e) within 1e-10.  Step sizes come from the
    device's pow() here, so the time grids agree to rounding, not bit for bit."""
    s = case[1]()
    nsteps = 400 if s.n == 2 else 200
    configure(ctx, s, False, None)
    o = Oracle(s, False, None)
    M = int(s.counts[:4].sum())
    t_o, h_o, att_o_total = 0.0, 0.05, 0
    for _ in range(nsteps):
        r_o, t_o, h_o, _, att_o, _ = o.step(integrator, t_o, h_o)
        assert r_o == 0
        att_o_total += att_o
    n0 = ctx.launch_count()
    rc, a, rec = ctx.run(integrator, 0.0, 0.05, nsteps, records=True)
    assert rc == 0 and a.steps == nsteps and a.stop_reason == capi.RUN_MAX_STEPS
    assert ctx.launch_count() - n0 == 1
    assert a.attempts == att_o_total, "identical accepted / rejected step counts"
    assert abs(a.time - t_o) <= 2e-2 * abs(t_o)
    assert np.all(np.diff(rec[:, 0]) > 0) and np.allclose(np.diff(rec[:, 0]), rec[1:, 1], rtol=1e-12)
    y_g, y_o = ctx.download(capi.Y0), o.array("y0")
    e_g, e_o = total_energy(y_g, s.mass, M), total_energy(y_o, s.mass, M)
    assert abs(e_g - e_o) <= 1e-10 * abs(e_o)
    a_g, ecc_g = orbital_elements_ae(y_g, s.mass)
    a_o, ecc_o = orbital_elements_ae(y_o, s.mass)
    assert np.max(np.abs(a_g - a_o) / np.abs(a_o)) <= 1e-10
    assert np.max(np.abs(ecc_g - ecc_o)) <= 1e-10


@pytest.mark.parametrize("integrator", [capi.RUNGE_KUTTA_FEHLBERG78, capi.DORMAND_PRINCE])
def test_run_adaptive_tracks_the_step_loop(ctx, integrator):
    """Same launch against this library's own single-step path up to the first snapshot of a length / output schedule: same
    number of steps and attempts, same stop reason, times within 1e-9 (device pow vs host pow in the step-size formula)."""
    s = synth.solar_system()
    kw = dict(length=4000.0, output=120.0)
    configure(ctx, s, False, None)
    ref = py_run(ctx, integrator, 0.0, 0.5, 600, **kw)
    y_ref = ctx.download(capi.Y0)
    configure(ctx, s, False, None)
    rc, a, rec = ctx.run(integrator, 0.0, 0.5, 600, records=True, **kw)
    assert rc == 0
    assert (a.steps, a.stop_reason, a.attempts) == (ref["steps"], ref["reason"], ref["attempts"])
    assert a.stop_reason == capi.RUN_SAVE and 10 < a.steps < 600
    assert np.allclose(rec[:, 0], ref["recs"][:, 0], rtol=1e-6, atol=0)        # a 1-ulp step-size difference moves the next errorMax by ~1e-6
    assert np.abs(ctx.download(capi.Y0) - y_ref).max() <= 1e-6 * np.abs(y_ref).max()


def test_run_stops_at_the_end_and_on_events(ctx):
    s = synth.solar_system()
    # end of the integration: |time| >= length, the last step clamped to it (Simulator.cpp:219,229-231)
    configure(ctx, s, False, None)
    rc, a, _ = ctx.run(capi.RUNGE_KUTTA4, 0.0, 1.0, 1000, length=25.5)
    assert rc == 0 and a.stop_reason == capi.RUN_END and a.steps == 26 and a.time == 25.5      # 25 full steps + one clamped to the length
    # hNext clamped to the length on the step before (Simulator.cpp:229-231): 25 full steps + one of 0.5
    configure(ctx, s, False, None)
    rc, a, rec = ctx.run(capi.RUNGE_KUTTA_FEHLBERG78, 0.0, 1.0, 1000, length=40.0, records=True)
    assert rc == 0 and a.stop_reason == capi.RUN_END and abs(a.time - 40.0) <= 1e-12 * 40.0
    # an ejection radius inside Neptune's orbit fires on the first step; the candidates are readable afterwards
    configure(ctx, s, False, None)
    rc, a, _ = ctx.run(capi.RUNGE_KUTTA4, 0.0, 1.0, 50, ejection=25.0)
    assert rc == 0 and a.stop_reason == capi.RUN_EVENT and a.steps == 1 and a.event_counts[0] == 1
    configure(ctx, s, False, None)
    ref = py_run(ctx, capi.RUNGE_KUTTA4, 0.0, 1.0, 50, ejection=25.0)
    assert ref["reason"] == capi.RUN_EVENT and ref["ev"] == list(a.event_counts)
    # a hit-centrum radius that Mercury crosses later: both loops stop on the same step
    configure(ctx, s, False, None)
    ref = py_run(ctx, capi.RUNGE_KUTTA4, 0.0, 0.5, 400, hit_centrum=0.32)
    configure(ctx, s, False, None)
    rc, a, _ = ctx.run(capi.RUNGE_KUTTA4, 0.0, 0.5, 400, hit_centrum=0.32)
    assert rc == 0 and (a.stop_reason, a.steps, a.time) == (ref["reason"], ref["steps"], ref["time"])
    assert a.stop_reason == capi.RUN_EVENT and 1 < a.steps < 400 and a.event_counts[1] == 1


@pytest.mark.parametrize("integrator", [capi.RUNGE_KUTTA4, capi.RUNGE_KUTTA_FEHLBERG78, capi.DORMAND_PRINCE])
def test_run_general_systems_is_bit_identical_to_the_step_loop(ctx, integrator):
    """Systems the one-warp kernel does not take (tracers, > 32 bodies) are stepped from the host inside sol_run with the
    very same drivers: bit-identical to sol_step + sol_detect_events + sol_flush_tiny."""
    s = synth.mixed([1, 2, 3, 10, 0, 60, 80], migration=False)
    kw = dict(length=1.0e9, output=9.0, flush_every=5, ejection=500.0)
    configure(ctx, s, False, None)
    ref = py_run(ctx, integrator, 0.0, 0.4, 40, **kw)
    y_ref = ctx.download(capi.Y0)
    configure(ctx, s, False, None)
    rc, a, rec = ctx.run(integrator, 0.0, 0.4, 40, records=True, **kw)
    assert rc == 0
    assert (a.steps, a.stop_reason, a.attempts, a.time, a.h_next, a.last_save) == (
        ref["steps"], ref["reason"], ref["attempts"], ref["time"], ref["h_next"], ref["last_save"])
    assert np.array_equal(rec, ref["recs"]) and np.array_equal(ctx.download(capi.Y0), y_ref)


def _close_pair(d):
    s = synth.mixed([1, 2, 0, 0, 0, 0, 0], migration=False)
    s.y0[2, :3] = s.y0[1, :3] + np.array([d, 0.0, 0.0])
    s.y0[2, 3:] = s.y0[1, 3:]
    return s


FAILURES = [
    # two giant planets 1e-8 au apart at t = 1e7 d: the step shrinks below ulp(t) -> RungeKuttaFehlberg78.cpp:116-122
    (capi.RUNGE_KUTTA_FEHLBERG78, 1.0e-8, 1.0e7, 100.0, "Stepsize-underflow occurred during Runge-Kutta-Fehlberg7(8) step!"),
    # 1e-5 au apart, h = 0.1 d: still errorMax ~ 2e-6 after 11 attempts -> DormandPrince.cpp:158-162
    (capi.DORMAND_PRINCE, 1.0e-5, 0.0, 0.1, "An error occurred during Prince-Dormand driver: iteration number exceeded maxIter!"),
]


@pytest.mark.parametrize("small_kernel", [1, 0], ids=["one-warp", "multi-launch"])
@pytest.mark.parametrize("case", FAILURES, ids=["rkf78-underflow", "dp-maxiter"])
def test_driver_failure_paths_match_the_reference(ctx, case, small_kernel):
    """The two failure returns of the adaptive drivers: the oracle (bit-exact restatement of the reference, which was
    checked to fail the same way) and the device both return 1 after the same number of attempts, with the reference's
    message - through sol_step and through sol_run."""
    integrator, d, t0, h0, msg = case
    s = _close_pair(d)
    o = Oracle(s, False, None)
    r_o, _, _, _, att_o, _ = o.step(integrator, t0, h0)
    assert r_o == 1
    configure(ctx, s, False, None)
    ctx.set_small_system_kernel(small_kernel)
    try:
        r_g, _, _, _, att_g, *_ = ctx.step(integrator, t0, h0)
        assert r_g == 1 and att_g == att_o
        assert ctx.last_error() == msg
        configure(ctx, s, False, None)
        rc, a, _ = ctx.run(integrator, t0, h0, 5)
        assert rc == 1 and a.stop_reason == capi.RUN_ERROR and ctx.last_error() == msg
    finally:
        ctx.set_small_system_kernel(1)


@pytest.mark.gpu
def test_fast_paths_of_sqrt_and_reciprocal_equal_the_library(ctx):
    """The persistent small-system kernel evaluates its own 1 / r^3 with straight-line copies of the fast paths of
    CUDA's double-precision sqrt and reciprocal (so that the scheduler can overlap them with the pair sums); outside the
    range in which both of the library's checks are known to pass it calls the library.  2^28 pseudo-random arguments
    over that whole range (random mantissas, runs of ones and zeros), compared bit for bit on the device with sqrt(x),
    1.0 / x and 1.0 / (x * sqrt(x))."""
    for seed in (1, 20260101):
        assert ctx.selftest_fast_paths(1 << 27, seed) == 0
