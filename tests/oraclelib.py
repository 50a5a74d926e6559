"""ctypes bindings for the two CHECKERS (test infrastructure only):

* ``Oracle``    - oracle/liboracle.so, the plain-C restatement (oracle/oracle.c)
* ``Reference`` - oracle/_ref/libref_harness.so, the compiled UNMODIFIED reference behind
                  oracle/ref_harness.cpp (present only where oracle/build_ref.sh has run)

Both expose the same Python surface so tests can swap them.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_SO = os.path.join(ROOT, "oracle", "liboracle.so")
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libref_harness.so")

EVAL_DRAG, EVAL_MIG1, EVAL_MIG2 = 1, 2, 4
EVAL_ALL = 7


class NebulaPod(C.Structure):
    """Layout shared by sol_nebula_pod (include/solaris_b200.h), oracle_nebula_pod and ref_nebula_pod."""
    _fields_ = [
        ("alpha", C.c_double), ("mean_molecular_weight", C.c_double), ("particle_diameter", C.c_double),
        ("decrease_type", C.c_int), ("_pad", C.c_int),
        ("time_scale", C.c_double), ("t0", C.c_double), ("t1", C.c_double),
        ("inner_edge", C.c_double),
        ("eta_c", C.c_double), ("eta_index", C.c_double),
        ("tau_c", C.c_double), ("tau_index", C.c_double),
        ("scale_height_c", C.c_double), ("scale_height_index", C.c_double),
        ("density_c", C.c_double), ("density_index", C.c_double),
        ("mean_free_path_c", C.c_double), ("mean_free_path_index", C.c_double),
    ]


def default_nebula() -> NebulaPod:
    """Values of a default-constructed GasComponent (Solaris/GasComponent.cpp:9-34), computed with the
    same expressions (checked against the compiled reference in tests/test_oracle_vs_reference.py)."""
    solar_to_kg = 1.98911e30
    au_to_m = 1.495978707e11
    gram_to_solar = 1.0 / (1.0e3 * solar_to_kg)
    meter_to_au = 1.0 / au_to_m
    gcm3 = gram_to_solar / ((1.0e-2 * meter_to_au) * (1.0e-2 * meter_to_au) * (1.0e-2 * meter_to_au))
    p = NebulaPod()
    p.alpha = 2.0e-3
    p.mean_molecular_weight = 2.3
    p.particle_diameter = 3.0e-10
    p.decrease_type = 0
    p.time_scale = p.t0 = p.t1 = 0.0
    p.inner_edge = 10.0 * (1.0 / 215.094)
    p.eta_c, p.eta_index = 0.0019, 0.5
    p.tau_c, p.tau_index = 2.0 / 3.0, 2.0
    p.scale_height_c, p.scale_height_index = 0.02, 1.25
    p.density_c, p.density_index = 1.4e-9 * gcm3, -2.75
    proton_cmu = 1.672621777e-27 * (1.0 / solar_to_kg)
    d_au = 3.0e-10 * meter_to_au
    p.mean_free_path_c = 2.3 * proton_cmu / (np.sqrt(2.0) * 3.14159265358979323846 * (d_au * d_au) * p.density_c)
    p.mean_free_path_index = 2.75
    return p


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def ensure_oracle_built():
    if not os.path.exists(ORACLE_SO) or os.path.getmtime(ORACLE_SO) < os.path.getmtime(os.path.join(ROOT, "oracle", "oracle.c")):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "liboracle.so"],
                              stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)


def reference_available() -> bool:
    return os.path.exists(REF_SO)


class _Base:
    prefix = ""
    lib = None

    def __init__(self, system, barycentric=False, nebula: NebulaPod | None = None, integrator=3):
        self.n = int(system.n)
        self.sys = system
        self._keep = [np.ascontiguousarray(system[k]) for k in (
            "y0", "mass", "radius", "density", "cD", "gammaStokes", "gammaEpstein", "migStopAt")]
        self._keepi = [np.ascontiguousarray(system[k], dtype=np.int32) for k in ("type", "migType", "id")]
        counts = (C.c_int * 7)(*[int(v) for v in system.counts])
        neb = C.byref(nebula) if nebula is not None else None
        args = [counts] + [_dp(a) for a in self._keep] + [_ip(a) for a in self._keepi] + [int(barycentric), neb]
        self.h = self._create(args, integrator)
        if not self.h:
            raise RuntimeError("create failed")

    def _f(self, name):
        return getattr(self.lib, self.prefix + name)

    def close(self):
        if self.h:
            self._f("destroy")(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def compute(self, t, y, flags=EVAL_ALL):
        y = np.ascontiguousarray(y, dtype=np.float64).reshape(-1)
        out = np.zeros(6 * self.n)
        r = self._f("compute")(self.h, C.c_double(t), _dp(y), _dp(out), C.c_uint(flags))
        if r != 0:
            raise RuntimeError("compute failed")
        return out.reshape(self.n, 6)

    def side(self):
        rm3 = np.zeros(self.n)
        idx = np.zeros(self.n, dtype=np.int32)
        dist = np.zeros(self.n)
        mig = np.zeros(self.n, dtype=np.int32)
        self._f("get_side")(self.h, _dp(rm3), _ip(idx), _dp(dist), _ip(mig))
        return rm3, idx, dist, mig

    def array(self, what):
        code = {"y0": 0, "y": 1, "accel": 2, "error": 3, "yscale": 4}[what]
        out = np.zeros(6 * self.n)
        self._f("get_array")(self.h, code, _dp(out))
        return out.reshape(self.n, 6)

    def set_y0(self, y0):
        y0 = np.ascontiguousarray(y0, dtype=np.float64).reshape(-1)
        self._f("set_y0")(self.h, _dp(y0))

    def flush_tiny(self):
        self._f("flush_tiny")(self.h)

    def remove_body(self, body_id):
        """Simulator::RemoveBody(bodyId) on this state ((f) row 3); returns the reference's return code."""
        f = self._f("remove_body")
        f.argtypes = [C.c_void_p, C.c_int]
        rc = f(self.h, int(body_id))
        if rc == 0:
            self.n -= 1
        return rc

    def params(self):
        n = self.n
        counts = np.zeros(7, dtype=np.int32)
        d = {k: np.zeros(n) for k in ("mass", "radius", "density", "cD", "gammaStokes", "gammaEpstein", "migStopAt")}
        i = {k: np.zeros(n, dtype=np.int32) for k in ("type", "migType", "id")}
        f = self._f("get_params")
        f.argtypes = [C.c_void_p] * 12
        f(self.h, counts.ctypes.data, *[d[k].ctypes.data for k in d], *[i[k].ctypes.data for k in i])
        d.update(i)
        d["counts"] = counts
        return d

    def event_records(self, ejection, hit_centrum, time, first_event_id=0):
        """Oracle only: the TwoBodyAffair.dat bytes of the ejection / hit-centrum scan (120 bytes per record)."""
        buf = np.zeros(120 * 2 * self.n, dtype=np.uint8)
        cnt = np.zeros(2, dtype=np.int32)
        f = self.lib.oracle_event_records
        f.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_int, C.c_void_p, C.POINTER(C.c_int)]
        n = f(self.h, ejection, hit_centrum, time, first_event_id, buf.ctypes.data, _ip(cnt))
        return buf[:120 * n].tobytes(), int(cnt[0]), int(cnt[1])

    def write_affairs(self, directory, filename, kinds, indices, time, first_event_id=0):
        """Reference only: its TwoBodyAffair constructor + SaveTwoBodyAffairs for the named events (scan order)."""
        kinds = np.ascontiguousarray(kinds, dtype=np.int32); indices = np.ascontiguousarray(indices, dtype=np.int32)
        f = self.lib.ref_write_affairs
        f.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_double, C.c_int]
        f.restype = None
        f(self.h, directory.encode(), filename.encode(), len(kinds), _ip(kinds), _ip(indices), time, first_event_id)

    def integrals(self):
        out = np.zeros(16)
        f = self._f("integrals")
        f.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
        f(self.h, _dp(out))
        return out


class Oracle(_Base):
    prefix = "oracle_"

    def __init__(self, *a, **k):
        ensure_oracle_built()
        if Oracle.lib is None:
            L = C.CDLL(ORACLE_SO)
            L.oracle_create.restype = C.c_void_p
            L.oracle_destroy.argtypes = [C.c_void_p]
            L.oracle_compute.argtypes = [C.c_void_p, C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_uint]
            L.oracle_get_side.argtypes = [C.c_void_p] + [C.c_void_p] * 4
            L.oracle_get_array.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_double)]
            L.oracle_set_y0.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
            L.oracle_flush_tiny.argtypes = [C.c_void_p]
            L.oracle_step.argtypes = [C.c_void_p, C.c_int] + [C.POINTER(C.c_double)] * 4
            L.oracle_gravity_rows.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_int, C.c_int, C.c_int]
            L.oracle_detect_events.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double] + [C.POINTER(C.c_int)] * 4
            L.oracle_time_gravity_rows.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.c_int, C.c_int, C.c_int, C.c_int]
            L.oracle_time_gravity_rows.restype = C.c_double
            L.oracle_gravity_rows_exact.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int), C.c_int, C.POINTER(C.c_double), C.c_int]
            L.oracle_gravity_row_quad.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.c_int, C.POINTER(C.c_double)]
            Oracle.lib = L
        super().__init__(*a, **k)

    def _create(self, args, integrator):
        return self.lib.oracle_create(*args)

    def step(self, integrator, time, h_next):
        t = C.c_double(time); hn = C.c_double(h_next); hd = C.c_double(0.0)
        info = (C.c_double * 2)()
        r = self.lib.oracle_step(self.h, integrator, C.byref(t), C.byref(hn), C.byref(hd), info)
        return r, t.value, hn.value, hd.value, int(info[0]), info[1]

    def gravity_rows(self, y, ib, ie, threads=1):
        y = np.ascontiguousarray(y, dtype=np.float64).reshape(-1)
        out = np.zeros(6 * self.n)
        r = self.lib.oracle_gravity_rows(self.h, _dp(y), _dp(out), ib, ie, threads)
        assert r == 0
        return out.reshape(self.n, 6)[ib:ie]

    def gravity_rows_exact(self, y, rows, threads=0):
        """Extended-precision (long double + compensated summation) value of the gravitational acceleration of the
        listed sinks, rounded to double: the 'true' sum the 1e-13 criterion is measured against at N ~ 10^6."""
        y = np.ascontiguousarray(y, dtype=np.float64).reshape(-1)
        rows = np.ascontiguousarray(rows, dtype=np.int32)
        out = np.zeros((len(rows), 3))
        r = self.lib.oracle_gravity_rows_exact(self.h, _dp(y), _ip(rows), len(rows), _dp(out), threads or (os.cpu_count() or 1))
        assert r == 0
        return out

    def gravity_row_quad(self, y, i):
        """The same sum in IEEE binary128 (libquadmath); slow, validates gravity_rows_exact."""
        y = np.ascontiguousarray(y, dtype=np.float64).reshape(-1)
        out = np.zeros(3)
        assert self.lib.oracle_gravity_row_quad(self.h, _dp(y), int(i), _dp(out)) == 0
        return out

    def time_gravity_rows(self, ib, ie, threads, reps):
        out = np.zeros(6 * self.n)
        return self.lib.oracle_time_gravity_rows(self.h, _dp(out), ib, ie, threads, reps)

    def detect_events(self, ejection, hit_centrum, collision_factor):
        ej = np.zeros(self.n, dtype=np.int32); hc = np.zeros(self.n, dtype=np.int32); co = np.zeros(self.n, dtype=np.int32)
        cnt = np.zeros(3, dtype=np.int32)
        self.lib.oracle_detect_events(self.h, ejection, hit_centrum, collision_factor, _ip(ej), _ip(hc), _ip(co), _ip(cnt))
        return ej[:cnt[0]].copy(), hc[:cnt[1]].copy(), co[:cnt[2]].copy()


class Reference(_Base):
    prefix = "ref_"

    @classmethod
    def _load(cls):
        if Reference.lib is None:
            L = C.CDLL(REF_SO)
            L.ref_create.restype = C.c_void_p
            L.ref_destroy.argtypes = [C.c_void_p]
            L.ref_compute.argtypes = [C.c_void_p, C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_uint]
            L.ref_get_side.argtypes = [C.c_void_p] + [C.c_void_p] * 4
            L.ref_get_array.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_double)]
            L.ref_set_y0.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
            L.ref_flush_tiny.argtypes = [C.c_void_p]
            L.ref_step.argtypes = [C.c_void_p, C.c_int] + [C.POINTER(C.c_double)] * 3
            L.ref_time_compute.argtypes = [C.c_void_p, C.c_double, C.c_int]
            L.ref_time_compute.restype = C.c_double
            L.ref_nebula_defaults.argtypes = [C.POINTER(NebulaPod)]
            Reference.lib = L

    def __init__(self, *a, **k):
        self._load()
        super().__init__(*a, **k)

    def _create(self, args, integrator):
        return self.lib.ref_create(*(args + [integrator]))

    def step(self, integrator, time, h_next):
        t = C.c_double(time); hn = C.c_double(h_next); hd = C.c_double(0.0)
        r = self.lib.ref_step(self.h, integrator, C.byref(t), C.byref(hn), C.byref(hd))
        return r, t.value, hn.value, hd.value, -1, float("nan")

    def time_compute(self, t, reps):
        return self.lib.ref_time_compute(self.h, t, reps)

    @staticmethod
    def nebula_defaults() -> NebulaPod:
        Reference._load()
        p = NebulaPod()
        Reference.lib.ref_nebula_defaults(C.byref(p))
        return p


# ---- (f) row 2: the Phases.dat snapshot writer (stateless helpers) ----
def oracle_pack_phases(time, y, ids):
    """oracle/oracle.c restatement of BinaryFileAdapter::SavePhases(BINARY): the bytes of one snapshot."""
    ensure_oracle_built()
    L = C.CDLL(ORACLE_SO)
    y = np.ascontiguousarray(y, dtype=np.float64).reshape(-1); ids = np.ascontiguousarray(ids, dtype=np.int32)
    n = len(ids)
    out = np.zeros(12 + 52 * n, dtype=np.uint8)
    L.oracle_pack_phases.restype = C.c_size_t
    L.oracle_pack_phases.argtypes = [C.c_double, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_int), C.c_void_p]
    k = L.oracle_pack_phases(time, n, _dp(y), _ip(ids), out.ctypes.data)
    assert k == out.size
    return out.tobytes()


def oracle_format_phases_text(time, y, ids):
    ensure_oracle_built()
    L = C.CDLL(ORACLE_SO)
    y = np.ascontiguousarray(y, dtype=np.float64).reshape(-1); ids = np.ascontiguousarray(ids, dtype=np.int32)
    n = len(ids)
    cap = 64 + 100 * n
    buf = C.create_string_buffer(cap)
    L.oracle_format_phases_text.restype = C.c_size_t
    L.oracle_format_phases_text.argtypes = [C.c_double, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_int), C.c_char_p, C.c_size_t]
    k = L.oracle_format_phases_text(time, n, _dp(y), _ip(ids), buf, cap)
    return buf.raw[:k]


def reference_save_phases(directory, filename, time, y, ids, text=False):
    """The compiled reference's own BinaryFileAdapter::SavePhases appending to directory/filename."""
    L = C.CDLL(REF_SO)
    y = np.ascontiguousarray(y, dtype=np.float64).reshape(-1).copy(); ids = np.ascontiguousarray(ids, dtype=np.int32).copy()
    L.ref_save_phases.argtypes = [C.c_char_p, C.c_char_p, C.c_double, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_int), C.c_int]
    L.ref_save_phases.restype = None
    L.ref_save_phases(directory.encode(), filename.encode(), time, len(ids), _dp(y), _ip(ids), 1 if text else 0)


# ---- (f) row 4: orbital elements -> phases (the loader's Kepler solves) ----
def _elements_to_phases(lib_path, fname, mu, el):
    L = C.CDLL(lib_path)
    mu = np.ascontiguousarray(mu, dtype=np.float64); el = np.ascontiguousarray(el, dtype=np.float64).reshape(-1, 6)
    out = np.zeros_like(el)
    f = getattr(L, fname)
    f.argtypes = [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double)]
    bad = f(len(mu), _dp(mu), _dp(el), _dp(out))
    return out, bad


def oracle_elements_to_phases(mu, el):
    """oracle/oracle.c restatement of Ephemeris::CalculatePhase; el rows = (a, e, incl, peri, node, M)."""
    ensure_oracle_built()
    return _elements_to_phases(ORACLE_SO, "oracle_elements_to_phases", mu, el)


def reference_elements_to_phases(mu, el):
    return _elements_to_phases(REF_SO, "ref_elements_to_phases", mu, el)
