"""Deterministic synthetic planetary systems for tests and bench.py.

Distributions follow the reference's own initial-condition generator
(src/Solaris.Initial.Cuda/main.cu:416-606, summarised in SURVEY.md §8d): star of 1 solar mass at the
origin, bodies on Keplerian orbits with a~U(5,6) AU (other ranges per config), e~U(0,0.1|0.2),
i=atan(0.05), angles ~U(0,2pi).  The reference seeds with time(0); here every variable draws from its
own counter-based Philox stream keyed by (seed, variable) so any N-prefix of a system is reproducible.

Arrays use the reference's host layout: bodies sorted by BodyType
(Solaris/Body.h:14-24), state AoS (x,y,z,vx,vy,vz) per body, body 0 = central body.
"""
from __future__ import annotations

import numpy as np

GAUSS = 1.720209895e-2                   # Solaris/Constants.h:28
GAUSS2 = 2.959122082855911025e-4         # Solaris/Constants.h:29
SOLAR_TO_JUPITER = 1.0473486e3
SOLAR_TO_SATURN = 3.497898e3
SOLAR_TO_EARTH = 3.3294605e5
SOLAR_TO_KG = 1.98911e30
AU_TO_METER = 1.495978707e11
GCM3_TO_SOLAR_AU3 = (1.0 / (1.0e3 * SOLAR_TO_KG)) / ((1.0e-2 * (1.0 / AU_TO_METER)) ** 3)

# BodyType (Solaris/Body.h:14-24)
CENTRAL, GIANT, ROCKY, PROTO, SUPERPL, PLANETESIMAL, TEST = 1, 2, 3, 4, 5, 6, 7
# MigrationType (Solaris/Body.h:35-39)
MIG_NO, MIG_I, MIG_II = 0, 1, 2
# IntegratorType (Solaris/IntegratorType.h:12-17)
DORMAND_PRINCE, RUNGE_KUTTA4, RUNGE_KUTTA_FEHLBERG78 = 0, 1, 3


def _rng(seed: int, var: int) -> np.random.Generator:
    return np.random.Generator(np.random.Philox(key=[seed & 0xFFFFFFFFFFFFFFFF, var]))


def kepler_E(M: np.ndarray, e: np.ndarray) -> np.ndarray:
    """Eccentric anomaly by Newton iteration (converged to 1e-15)."""
    E = M + e * np.sin(M)
    if E.size == 0:
        return E
    for _ in range(50):
        dE = (E - e * np.sin(E) - M) / (1.0 - e * np.cos(E))
        E = E - dE
        if np.max(np.abs(dE)) < 1e-15:
            break
    return E


def elements_to_phase(mu, a, e, inc, peri, node, M):
    """Orbital elements -> astrocentric phase (x,y,z,vx,vy,vz); vectorised."""
    mu, a, e, inc, peri, node, M = (np.asarray(v, dtype=np.float64) for v in (mu, a, e, inc, peri, node, M))
    E = kepler_E(M, e)
    cosE, sinE = np.cos(E), np.sin(E)
    # position / velocity in the orbital plane
    xp = a * (cosE - e)
    yp = a * np.sqrt(1.0 - e * e) * sinE
    r = a * (1.0 - e * cosE)
    k = np.sqrt(mu * a) / r
    vxp = -k * sinE
    vyp = k * np.sqrt(1.0 - e * e) * cosE
    cw, sw = np.cos(peri), np.sin(peri)
    cO, sO = np.cos(node), np.sin(node)
    ci, si = np.cos(inc), np.sin(inc)
    P = np.stack([cw * cO - sw * sO * ci, cw * sO + sw * cO * ci, sw * si], axis=-1)
    Q = np.stack([-sw * cO - cw * sO * ci, -sw * sO + cw * cO * ci, cw * si], axis=-1)
    pos = xp[..., None] * P + yp[..., None] * Q
    vel = vxp[..., None] * P + vyp[..., None] * Q
    return np.concatenate([pos, vel], axis=-1)


class System(dict):
    """dict with attribute access; keys: counts,y0,mass,radius,density,cD,gammaStokes,gammaEpstein,
    migStopAt,type,migType,id,n"""
    __getattr__ = dict.__getitem__
    __setattr__ = dict.__setitem__


def _finish(counts, y0, mass, radius, density, cD, migStopAt, migType) -> System:
    counts = np.asarray(counts, dtype=np.int32)
    n = int(counts.sum())
    types = np.repeat(np.arange(1, 8, dtype=np.int32), counts)
    mass = np.ascontiguousarray(mass, dtype=np.float64)
    radius = np.ascontiguousarray(radius, dtype=np.float64)
    density = np.ascontiguousarray(density, dtype=np.float64)
    cD = np.ascontiguousarray(cD, dtype=np.float64)
    # Simulator::BodyListToBodyData (Solaris/Simulator.cpp:543-568): gammas only for radius > 0,
    # gammaStokes only for cd > 0, test particles carry zeros.
    gE = np.zeros(n)
    gS = np.zeros(n)
    ok = (radius > 0) & (types != TEST)
    gE[ok] = 1.0 / (density[ok] * radius[ok])
    okS = ok & (cD > 0)
    gS[okS] = (3.0 / 8.0) * cD[okS] / (density[okS] * radius[okS])
    tp = types == TEST
    for arr in (mass, radius, density, cD):
        arr[tp] = 0.0
    return System(
        counts=counts, n=n,
        y0=np.ascontiguousarray(y0, dtype=np.float64).reshape(n, 6),
        mass=mass, radius=radius, density=density, cD=cD, gammaStokes=gS, gammaEpstein=gE,
        migStopAt=np.ascontiguousarray(migStopAt, dtype=np.float64),
        type=types, migType=np.ascontiguousarray(migType, dtype=np.int32),
        id=np.arange(1, n + 1, dtype=np.int32),
    )


def _disk_elements(seed, n, a_rng=(5.0, 6.0), e_max=0.1, inc=None):
    a = _rng(seed, 1).uniform(a_rng[0], a_rng[1], n)
    e = _rng(seed, 2).uniform(0.0, e_max, n)
    if inc is None:
        i = np.full(n, np.arctan(0.05))
    else:
        i = _rng(seed, 3).uniform(inc[0], inc[1], n)
    w = _rng(seed, 4).uniform(0.0, 2 * np.pi, n)
    O = _rng(seed, 5).uniform(0.0, 2 * np.pi, n)
    M = _rng(seed, 6).uniform(0.0, 2 * np.pi, n)
    return a, e, i, w, O, M


def massive_disk(n: int, seed: int = 20240601 + 5, migration: bool = False, body_type: int = PROTO) -> System:
    """Star + (n-1) self-gravitating bodies (configs C5 / H): m~U(0.001,0.1) M_earth,
    a~U(5,6), e~U(0,0.1).  With migration=True every body carries Migration type I, stopAt 0.4."""
    nb = n - 1
    a, e, i, w, O, M = _disk_elements(seed, nb)
    m = _rng(seed, 7).uniform(0.001, 0.1, nb) / SOLAR_TO_EARTH
    mu = GAUSS2 * (1.0 + m)
    ph = elements_to_phase(mu, a, e, i, w, O, M)
    y0 = np.vstack([np.zeros((1, 6)), ph])
    mass = np.concatenate([[1.0], m])
    dens = np.concatenate([[0.0], _rng(seed, 8).uniform(1.0, 2.0, nb) * GCM3_TO_SOLAR_AU3])
    radius = np.zeros(n)
    radius[1:] = (3.0 / (4.0 * np.pi) * mass[1:] / dens[1:]) ** (1.0 / 3.0)
    counts = [1, 0, 0, 0, 0, 0, 0]
    counts[body_type - 1] += nb
    migType = np.zeros(n, dtype=np.int32)
    stop = np.zeros(n)
    if migration:
        migType[1:] = MIG_I
        stop[1:] = 0.4
    return _finish(counts, y0, mass, radius, dens, np.zeros(n), stop, migType)


def _jupiter_saturn(with_saturn=True):
    deg = np.pi / 180.0
    mj = 1.0 / SOLAR_TO_JUPITER
    # Jupiter elements of TestCases/SunJupiter/SunJupiter.xml:27
    el = [(5.20336301, 0.04839266, 1.3053 * deg, 274.1977 * deg, 100.55615 * deg, 19.65053 * deg, mj)]
    if with_saturn:
        ms = 1.0 / SOLAR_TO_SATURN
        el.append((9.53707032, 0.0541506, 2.48446 * deg, 338.7169 * deg, 113.71504 * deg, 317.51238 * deg, ms))
    a, e, i, w, O, M, m = (np.array(v) for v in zip(*el))
    ph = elements_to_phase(GAUSS2 * (1.0 + m), a, e, i, w, O, M)
    return ph, m


def sun_jupiter() -> System:
    """Config C1: Sun + Jupiter with the elements of TestCases/SunJupiter/SunJupiter.xml:27."""
    ph, m = _jupiter_saturn(False)
    y0 = np.vstack([np.zeros((1, 6)), ph])
    z = np.zeros(2)
    return _finish([1, 1, 0, 0, 0, 0, 0], y0, np.concatenate([[1.0], m]), z.copy(), z.copy(), z.copy(), z.copy(),
                   np.zeros(2, dtype=np.int32))


def solar_system() -> System:
    """Config C2: Sun + 8 planets in the reference's body order (giants first: Jupiter, Saturn, Uranus,
    Neptune, then the rocky planets Earth, Mercury, Venus, Mars = order of
    TestCases/SolarSystemWithBalint/SS.data).  J2000 mean elements (public ephemeris values), masses
    from the reference's Solar-to-planet ratios (Solaris/Constants.h:34-42)."""
    deg = np.pi / 180.0
    #        a [AU]       e           i [deg]   peri [deg]  node [deg]  M [deg]    Msun/m
    el = [
        (5.20336301, 0.04839266, 1.30530, 274.19770, 100.55615, 19.65053, 1.0473486e3),   # Jupiter
        (9.53707032, 0.05415060, 2.48446, 338.71690, 113.71504, 317.51238, 3.497898e3),   # Saturn
        (19.19126393, 0.04716771, 0.76986, 96.73436, 74.22988, 142.26794, 2.290298e4),    # Uranus
        (30.06896348, 0.00858587, 1.76917, 273.24966, 131.72169, 259.90868, 1.941224e4),  # Neptune
        (1.00000011, 0.01671022, 0.00005, 114.20783, 348.73936, 357.51716, 3.3294605e5),  # Earth
        (0.38709893, 0.20563069, 7.00487, 29.12478, 48.33167, 174.79439, 6.0236e6),       # Mercury
        (0.72333199, 0.00677323, 3.39471, 54.85229, 76.68069, 50.44675, 4.0852371e5),     # Venus
        (1.52366231, 0.09341233, 1.85061, 286.46230, 49.57854, 19.41248, 3.098708e6),     # Mars
    ]
    a, e, i, w, O, M, inv = (np.array(v) for v in zip(*el))
    m = 1.0 / inv
    ph = elements_to_phase(GAUSS2 * (1.0 + m), a, e, i * deg, w * deg, O * deg, M * deg)
    n = 9
    y0 = np.vstack([np.zeros((1, 6)), ph])
    z = np.zeros(n)
    return _finish([1, 4, 4, 0, 0, 0, 0], y0, np.concatenate([[1.0], m]), z.copy(), z.copy(), z.copy(), z.copy(),
                   np.zeros(n, dtype=np.int32))


def trojans(n_test: int, seed: int = 20240601 + 4) -> System:
    """Sun + Jupiter + Saturn + n_test test particles around L4/L5 (config C4)."""
    ph_p, m_p = _jupiter_saturn(True)
    deg = np.pi / 180.0
    a = _rng(seed, 1).uniform(5.05, 5.35, n_test)
    e = _rng(seed, 2).uniform(0.0, 0.15, n_test)
    i = _rng(seed, 3).uniform(0.0, 0.4, n_test)
    w = _rng(seed, 4).uniform(0.0, 2 * np.pi, n_test)
    O = _rng(seed, 5).uniform(0.0, 2 * np.pi, n_test)
    lam_j = (274.1977 + 100.55615 + 19.65053) * deg
    side = np.where(_rng(seed, 6).random(n_test) < 0.5, 1.0, -1.0)
    lam = lam_j + side * 60.0 * deg + _rng(seed, 7).normal(0.0, 10.0 * deg, n_test)
    M = np.mod(lam - w - O, 2 * np.pi)
    ph = elements_to_phase(np.full(n_test, GAUSS2), a, e, i, w, O, M)
    n = 3 + n_test
    y0 = np.vstack([np.zeros((1, 6)), ph_p, ph])
    mass = np.concatenate([[1.0], m_p, np.zeros(n_test)])
    z = np.zeros(n)
    return _finish([1, 2, 0, 0, 0, 0, n_test], y0, mass, z.copy(), z.copy(), z.copy(), z.copy(), np.zeros(n, dtype=np.int32))


def planetesimal_drag(n_pl: int, seed: int = 20240601 + 3) -> System:
    """Sun + Jupiter + n_pl planetesimals that feel gas drag (config C3): a~U(1,4), e~U(0,0.2),
    rho~U(1,2) g/cm3, R~U(5,15) km, Cd~U(0.5,4)."""
    ph_p, m_p = _jupiter_saturn(False)
    a, e, i, w, O, M = _disk_elements(seed, n_pl, a_rng=(1.0, 4.0), e_max=0.2)
    dens = _rng(seed, 8).uniform(1.0, 2.0, n_pl) * GCM3_TO_SOLAR_AU3
    R = _rng(seed, 9).uniform(5.0, 15.0, n_pl) * 1.0e3 / AU_TO_METER
    cd = _rng(seed, 10).uniform(0.5, 4.0, n_pl)
    m = dens * (4.0 / 3.0 * np.pi * R ** 3)
    ph = elements_to_phase(GAUSS2 * (1.0 + m), a, e, i, w, O, M)
    n = 2 + n_pl
    y0 = np.vstack([np.zeros((1, 6)), ph_p, ph])
    mass = np.concatenate([[1.0], m_p, m])
    return _finish([1, 1, 0, 0, 0, n_pl, 0], y0, mass,
                   np.concatenate([[0.0, 0.0], R]), np.concatenate([[0.0, 0.0], dens]),
                   np.concatenate([[0.0, 0.0], cd]), np.zeros(n), np.zeros(n, dtype=np.int32))


def mixed(counts, seed: int = 20240601 + 9, migration: bool = True, a_rng=(1.0, 6.0)) -> System:
    """Every body type at once (parity tests): counts = [1, giant, rocky, proto, superpl, pl, test]."""
    counts = list(counts)
    assert counts[0] == 1
    n = int(sum(counts))
    nb = n - 1
    types = np.repeat(np.arange(1, 8), counts)[1:]
    a, e, i, w, O, M = _disk_elements(seed, nb, a_rng=a_rng, e_max=0.2)
    u = _rng(seed, 7).random(nb)
    m = np.zeros(nb)
    m[types == GIANT] = (0.1 + 9.9 * u[types == GIANT]) / SOLAR_TO_JUPITER
    m[types == ROCKY] = (0.1 + 9.9 * u[types == ROCKY]) / SOLAR_TO_EARTH
    m[types == PROTO] = (0.001 + 0.099 * u[types == PROTO]) / SOLAR_TO_EARTH
    dens = _rng(seed, 8).uniform(1.0, 2.0, nb) * GCM3_TO_SOLAR_AU3
    small = (types == SUPERPL) | (types == PLANETESIMAL)
    # radii spanning the Epstein / transition / Stokes regimes at a few AU (mean free path ~ 1e-11..1e-9 AU)
    R = np.zeros(nb)
    R[small] = 10.0 ** _rng(seed, 9).uniform(-13.0, -7.5, int(small.sum()))
    big = ~small & (types != TEST)
    R[big] = (3.0 / (4.0 * np.pi) * m[big] / dens[big]) ** (1.0 / 3.0)
    m[types == PLANETESIMAL] = dens[types == PLANETESIMAL] * (4.0 / 3.0 * np.pi * R[types == PLANETESIMAL] ** 3)
    m[types == SUPERPL] = 1.0e-9 * (0.5 + u[types == SUPERPL])
    cd = np.where(small, _rng(seed, 10).uniform(0.5, 4.0, nb), 0.0)
    mu = GAUSS2 * (1.0 + m)
    ph = elements_to_phase(mu, a, e, i, w, O, M)
    y0 = np.vstack([np.zeros((1, 6)), ph])
    migType = np.zeros(n, dtype=np.int32)
    stop = np.zeros(n)
    if migration:
        tt = np.concatenate([[CENTRAL], types])
        sel1 = (tt == ROCKY) | (tt == PROTO)
        migType[sel1] = MIG_I
        migType[tt == GIANT] = MIG_II
        # a few bodies already inside their stop radius so the "migType -> No" side effect is covered
        r = np.sqrt((y0[:, :3] ** 2).sum(axis=1))
        stop[:] = 0.4
        flip = (_rng(seed, 11).random(n) < 0.15) & (migType != MIG_NO)
        stop[flip] = r[flip] * 1.5
    return _finish(counts, y0, np.concatenate([[1.0], m]), np.concatenate([[0.0], R]),
                   np.concatenate([[0.0], dens]), np.concatenate([[0.0], cd]), stop, migType)


def to_barycentric(sys_: System) -> System:
    """Shift phases to the barycentre of the massive bodies (Calculate::PhaseWithRespectToBC,
    Solaris/Simulator.cpp:577-581)."""
    out = System({k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in sys_.items()})
    M = int(out.counts[:4].sum())
    mt = out.mass[:M].sum()
    bc = (out.mass[:M, None] * out.y0[:M]).sum(axis=0) / mt
    out.y0 = out.y0 - bc[None, :]
    return out


def pairs_per_eval(counts, barycentric: bool) -> int:
    """Ordered (sink, source) pairs per force evaluation (SURVEY.md §8d 'Unit of work')."""
    c = [int(v) for v in counts]
    n = sum(c)
    M = c[0] + c[1] + c[2] + c[3]
    s = c[4]
    if barycentric:
        return n * M - M
    pairs = 0
    # massive sinks 1..M-1: sources 1..M+s-1 except self
    pairs += (M - 1) * (M + s - 1 - 1)
    # non-massive sinks M..n-1: sources 1..M-1
    pairs += (n - M) * (M - 1)
    return pairs
