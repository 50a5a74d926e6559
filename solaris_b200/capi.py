"""ctypes binding of the C-ABI in include/solaris_b200.h (libsolaris_b200.so).

This is only a thin convenience layer for tests and bench.py: every method is one C-ABI call with
numpy host buffers.  There is no CPU fallback - if the shared library is missing or no CUDA device
is usable, construction raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SOLARIS_B200_LIB") or os.path.join(HERE, "libsolaris_b200.so")   # (A/B builds: tools/build_variant.py)

EVAL_GAS_DRAG, EVAL_MIG_TYPE1, EVAL_MIG_TYPE2, EVAL_ALL = 1, 2, 4, 7
DORMAND_PRINCE, RUNGE_KUTTA4, RUNGE_KUTTA_FEHLBERG78 = 0, 1, 3

Y0, Y, ACCEL, YSCALE, RM3, NN_INDEX, NN_DISTANCE, MIGTYPE, MASS, RADIUS = range(10)
DENSITY, CD, GAMMA_STOKES, GAMMA_EPSTEIN, MIGSTOPAT, TYPE, ID = range(13, 20)
ACCEL_GASDRAG, ACCEL_MIGTYPE1, ACCEL_MIGTYPE2 = 10, 11, 12

# every symbol declared in include/solaris_b200.h (tests check the library exports all of them)
EXPORTS = [
    "sol_create", "sol_create_multi", "sol_destroy", "sol_last_error", "sol_set_stream", "sol_set_bodies", "sol_set_frame",
    "sol_set_nebula", "sol_set_nn_tracking", "sol_set_pair_algorithm", "sol_set_small_system_kernel", "sol_set_tracer_kernel", "sol_set_graph_mode", "sol_compute", "sol_compute_device", "sol_step", "sol_run",
    "sol_detect_events", "sol_event_indices", "sol_event_records", "sol_integrals", "sol_pack_phases", "sol_write_phases", "sol_remove_bodies", "sol_patch_body", "sol_elements_to_phases", "sol_download", "sol_upload", "sol_flush_tiny",
    "sol_body_count", "sol_nccl_unique_id", "sol_dist_init", "sol_shard_of", "sol_sym_round_pair", "sol_sym_rounds_of_rank", "sol_sym_work_of_rank", "sol_plan_pairs", "sol_shard_range", "sol_gather_state",
    "sol_time_gravity_kernel", "sol_measure_fp64_peak", "sol_selftest_fast_paths", "sol_launch_count", "sol_profile_enable",
    "sol_profile_read",
]


class NebulaPod(C.Structure):
    """sol_nebula_pod (include/solaris_b200.h)."""
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
    """A default-constructed GasComponent (Solaris/GasComponent.cpp:9-34), same expressions."""
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


class RunArgs(C.Structure):
    """sol_run_args (include/solaris_b200.h)."""
    _fields_ = [
        ("integrator", C.c_int), ("max_steps", C.c_int),
        ("time", C.c_double), ("h_next", C.c_double), ("h_did", C.c_double),
        ("millenium_days", C.c_double), ("length", C.c_double), ("output", C.c_double), ("last_save", C.c_double),
        ("ejection", C.c_double), ("hit_centrum", C.c_double), ("collision_factor", C.c_double),
        ("step_counter", C.c_longlong), ("flush_every", C.c_int), ("flush_threshold", C.c_double),
        ("steps", C.c_int), ("stop_reason", C.c_int), ("event_counts", C.c_int * 3),
        ("attempts", C.c_longlong), ("err_max", C.c_double), ("records", C.POINTER(C.c_double)),
    ]


RUN_MAX_STEPS, RUN_END, RUN_SAVE, RUN_EVENT, RUN_ERROR = range(5)

_lib = None


def load_library() -> C.CDLL:
    """Loads libsolaris_b200.so; raises if it has not been built (python -m solaris_b200.build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: build it with `python -m solaris_b200.build` "
                           "(solaris_b200 has no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    dp, ip, vp = C.POINTER(C.c_double), C.POINTER(C.c_int), C.c_void_p
    L.sol_create.argtypes = [C.c_int, C.POINTER(vp)]
    L.sol_create_multi.argtypes = [C.c_int, C.POINTER(vp)]
    L.sol_destroy.argtypes = [vp]
    L.sol_destroy.restype = None
    L.sol_last_error.argtypes = [vp]
    L.sol_last_error.restype = C.c_char_p
    L.sol_set_stream.argtypes = [vp, vp]
    L.sol_set_bodies.argtypes = [vp, ip] + [dp] * 8 + [ip] * 3
    L.sol_set_frame.argtypes = [vp, C.c_int]
    L.sol_set_nebula.argtypes = [vp, C.POINTER(NebulaPod)]
    L.sol_set_nn_tracking.argtypes = [vp, C.c_int]
    L.sol_set_pair_algorithm.argtypes = [vp, C.c_int]
    L.sol_set_small_system_kernel.argtypes = [vp, C.c_int]
    L.sol_set_tracer_kernel.argtypes = [vp, C.c_int]
    L.sol_set_graph_mode.argtypes = [vp, C.c_int]
    L.sol_compute.argtypes = [vp, C.c_double, vp, vp, C.c_uint]
    L.sol_compute_device.argtypes = [vp, C.c_double, C.c_uint]
    L.sol_step.argtypes = [vp, C.c_int, dp, dp, dp, dp]
    L.sol_run.argtypes = [vp, C.POINTER(RunArgs)]
    L.sol_detect_events.argtypes = [vp, C.c_double, C.c_double, C.c_double, ip]
    L.sol_event_indices.argtypes = [vp, C.c_int, ip, C.c_int, ip]
    L.sol_integrals.argtypes = [vp, dp]
    L.sol_event_records.argtypes = [vp, C.c_double, C.c_int, C.c_void_p, C.c_int, C.POINTER(C.c_int)]
    L.sol_download.argtypes = [vp, C.c_int, vp]
    L.sol_upload.argtypes = [vp, C.c_int, vp]
    L.sol_flush_tiny.argtypes = [vp, C.c_double]
    L.sol_pack_phases.argtypes = [vp, C.c_double, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
    L.sol_write_phases.argtypes = [vp, C.c_char_p, C.c_double]
    L.sol_remove_bodies.argtypes = [vp, C.POINTER(C.c_int), C.c_int]
    L.sol_elements_to_phases.argtypes = [vp, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_int)]
    L.sol_patch_body.argtypes = [vp, C.c_int, C.POINTER(C.c_double), C.c_double, C.c_double, C.c_double]
    L.sol_body_count.argtypes = [vp]
    L.sol_nccl_unique_id.argtypes = [vp]
    L.sol_dist_init.argtypes = [vp, C.c_int, C.c_int, vp]
    L.sol_shard_of.argtypes = [C.c_int, C.c_int, C.c_int, ip, ip]
    L.sol_sym_round_pair.argtypes = [C.c_int, C.c_int, C.c_int, ip]
    L.sol_sym_rounds_of_rank.argtypes = [C.c_int, C.c_int, C.c_int, ip, ip]
    L.sol_sym_work_of_rank.argtypes = [C.c_int, C.c_int, C.c_int, ip]
    L.sol_plan_pairs.argtypes = [C.c_int, C.c_int, C.c_int, ip]
    L.sol_shard_range.argtypes = [vp, ip, ip]
    L.sol_gather_state.argtypes = [vp]
    L.sol_time_gravity_kernel.argtypes = [vp, C.c_int, C.POINTER(C.c_float), dp]
    L.sol_measure_fp64_peak.argtypes = [vp, dp]
    L.sol_selftest_fast_paths.argtypes = [vp, C.c_ulonglong, C.c_longlong, C.POINTER(C.c_ulonglong)]
    L.sol_launch_count.argtypes = [vp]
    L.sol_launch_count.restype = C.c_longlong
    L.sol_profile_enable.argtypes = [vp, C.c_int]
    L.sol_profile_read.argtypes = [vp, dp, C.POINTER(C.c_longlong), C.c_int]
    _lib = L
    return L


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


class SolarisError(RuntimeError):
    pass


def shard_of(n: int, nranks: int, rank: int):
    """Sink range [lo, hi) of `rank` (sol_shard_of; usable without a GPU)."""
    lo, hi = C.c_int(0), C.c_int(0)
    if load_library().sol_shard_of(n, nranks, rank, C.byref(lo), C.byref(hi)) != 0:
        raise SolarisError("sol_shard_of: bad arguments")
    return lo.value, hi.value


class Context:
    """One sol_ctx: the device-resident system of one process / one GPU."""

    def __init__(self, device: int = 0, n_gpus: int = 0):
        """device: one context on that GPU.  n_gpus >= 1: one handle over the first n_gpus devices of this process
        (sol_create_multi: sinks sharded over the devices, one worker thread each)."""
        self.lib = load_library()
        h = C.c_void_p()
        rc = self.lib.sol_create_multi(n_gpus, C.byref(h)) if n_gpus >= 1 else self.lib.sol_create(device, C.byref(h))
        if rc != 0:
            raise SolarisError(self.lib.sol_last_error(None).decode())
        self.h = h
        self.n = 0
        self.counts = None

    def close(self):
        if getattr(self, "h", None):
            self.lib.sol_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, r):
        if r != 0:
            raise SolarisError(self.lib.sol_last_error(self.h).decode())

    # ---- configuration ----
    def set_stream(self, cuda_stream_ptr: int):
        self._check(self.lib.sol_set_stream(self.h, C.c_void_p(cuda_stream_ptr)))

    def set_frame(self, barycentric: bool):
        self._check(self.lib.sol_set_frame(self.h, int(barycentric)))

    def set_nebula(self, nebula):
        if nebula is None:
            self._check(self.lib.sol_set_nebula(self.h, None))
        else:
            pod = NebulaPod.from_buffer_copy(bytes(nebula))
            self._check(self.lib.sol_set_nebula(self.h, C.byref(pod)))

    def set_nn_tracking(self, mode: int):
        self._check(self.lib.sol_set_nn_tracking(self.h, mode))

    def set_pair_algorithm(self, mode: int):
        self._check(self.lib.sol_set_pair_algorithm(self.h, mode))

    def set_small_system_kernel(self, on):
        self._check(self.lib.sol_set_small_system_kernel(self.h, int(on)))

    def set_graph_mode(self, mode: int):
        """0: every launch from the host, 1 (default): CUDA-graph replay, 2: one cooperative kernel per segment."""
        self._check(self.lib.sol_set_graph_mode(self.h, int(mode)))

    def set_tracer_kernel(self, on: bool):
        self._check(self.lib.sol_set_tracer_kernel(self.h, int(on)))

    def set_bodies(self, system):
        counts = np.ascontiguousarray(system["counts"], dtype=np.int32)
        d = [np.ascontiguousarray(system[k], dtype=np.float64) for k in (
            "y0", "mass", "radius", "density", "cD", "gammaStokes", "gammaEpstein", "migStopAt")]
        i = [np.ascontiguousarray(system[k], dtype=np.int32) for k in ("type", "migType", "id")]
        self._check(self.lib.sol_set_bodies(self.h, _ip(counts), *[_dp(a) for a in d], *[_ip(a) for a in i]))
        self.n = int(counts.sum())
        self.counts = counts

    # ---- seam B ----
    def compute(self, t: float, y: np.ndarray, flags: int = EVAL_ALL, out: np.ndarray | None = None) -> np.ndarray:
        y = np.ascontiguousarray(y, dtype=np.float64)
        if out is None:
            out = np.empty((self.n, 6))
        self._check(self.lib.sol_compute(self.h, t, y.ctypes.data, out.ctypes.data, flags))
        return out

    def compute_ptr(self, t: float, y_ptr: int, out_ptr: int, flags: int = EVAL_ALL):
        """Same call with raw host pointers (e.g. pinned torch tensors)."""
        self._check(self.lib.sol_compute(self.h, t, C.c_void_p(y_ptr), C.c_void_p(out_ptr), flags))

    def compute_device(self, t: float, flags: int = EVAL_ALL):
        self._check(self.lib.sol_compute_device(self.h, t, flags))

    # ---- seam A ----
    def step(self, integrator: int, time: float, h_next: float):
        """Returns (rc, time, hNext, hDid, attempts, errorMax, evals, pairs); rc 1 = driver failure."""
        t = C.c_double(time); hn = C.c_double(h_next); hd = C.c_double(0.0)
        info = (C.c_double * 4)()
        r = self.lib.sol_step(self.h, integrator, C.byref(t), C.byref(hn), C.byref(hd), info)
        return r, t.value, hn.value, hd.value, int(info[0]), info[1], info[2], info[3]

    def run(self, integrator: int, time: float, h_next: float, max_steps: int, length: float = 1.0e300, output: float = 1.0e300,
            last_save: float = 0.0, millenium_days: float = 0.0, ejection: float = 0.0, hit_centrum: float = 0.0,
            collision_factor: float = 0.0, step_counter: int = 0, flush_every: int = 100, flush_threshold: float = 1.0e-50,
            records: bool = False):
        """sol_run: many Driver steps in one call.  Returns (rc, RunArgs, records or None); records[k] = (time, hDid, hNext, trial h)."""
        a = RunArgs()
        a.integrator = integrator; a.max_steps = max_steps; a.time = time; a.h_next = h_next
        a.millenium_days = millenium_days; a.length = length; a.output = output; a.last_save = last_save
        a.ejection = ejection; a.hit_centrum = hit_centrum; a.collision_factor = collision_factor
        a.step_counter = step_counter; a.flush_every = flush_every; a.flush_threshold = flush_threshold
        rec = None
        if records:
            rec = np.zeros((max_steps, 4))
            a.records = _dp(rec)
        rc = self.lib.sol_run(self.h, C.byref(a))
        return rc, a, (rec[:a.steps] if rec is not None else None)

    def last_error(self) -> str:
        return self.lib.sol_last_error(self.h).decode()

    # ---- events ----
    def detect_events(self, ejection: float, hit_centrum: float, collision_factor: float):
        cnt = np.zeros(3, dtype=np.int32)
        self._check(self.lib.sol_detect_events(self.h, ejection, hit_centrum, collision_factor, _ip(cnt)))
        out = []
        for kind in range(3):
            idx = np.zeros(max(int(cnt[kind]), 1), dtype=np.int32)
            m = C.c_int(0)
            self._check(self.lib.sol_event_indices(self.h, kind, _ip(idx), int(cnt[kind]), C.byref(m)))
            out.append(idx[:m.value].copy())      # sharded contexts hold the candidates of their own sinks only
        return out

    def event_records(self, time: float, first_event_id: int = 0) -> bytes:
        """TwoBodyAffair.dat bytes (120 per record) for the ejections / hit centrums of the last detect_events()."""
        m = C.c_int(0)
        self._check(self.lib.sol_event_records(self.h, time, first_event_id, None, 0, C.byref(m)))
        buf = np.zeros(120 * max(m.value, 1), dtype=np.uint8)
        self._check(self.lib.sol_event_records(self.h, time, first_event_id, buf.ctypes.data, m.value, C.byref(m)))
        return buf[:120 * m.value].tobytes()

    def integrals(self) -> np.ndarray:
        out = np.zeros(16)
        self._check(self.lib.sol_integrals(self.h, _dp(out)))
        return out

    def remove_bodies(self, indices) -> None:
        """Simulator::RemoveBody for the bodies at these current indices, on the device."""
        idx = np.ascontiguousarray(indices, dtype=np.int32)
        self._check(self.lib.sol_remove_bodies(self.h, _ip(idx), len(idx)))
        self.n = int(self.lib.sol_body_count(self.h))
        self.counts = np.bincount(self.download(TYPE), minlength=8)[1:8].astype(np.int32)

    def patch_body(self, index: int, y0, mass: float, radius: float, density: float) -> None:
        y = np.ascontiguousarray(y0, dtype=np.float64)
        self._check(self.lib.sol_patch_body(self.h, int(index), _dp(y), mass, radius, density))

    def elements_to_phases(self, mu, elements):
        """Ephemeris::CalculatePhase for a batch; returns (phases, number of non-converged bodies)."""
        mu = np.ascontiguousarray(mu, dtype=np.float64)
        el = np.ascontiguousarray(elements, dtype=np.float64).reshape(-1, 6)
        out = np.zeros_like(el)
        bad = C.c_int(0)
        rc = self.lib.sol_elements_to_phases(self.h, len(mu), _dp(mu), _dp(el), _dp(out), C.byref(bad))
        if rc != 0 and bad.value == 0:
            self._check(rc)
        return out, bad.value

    def pack_phases(self, time: float) -> bytes:
        """The Phases.dat record of the resident state (BinaryFileAdapter::SavePhases, BINARY)."""
        nb = C.c_size_t(0)
        self._check(self.lib.sol_pack_phases(self.h, time, None, 0, C.byref(nb)))
        buf = np.zeros(nb.value, dtype=np.uint8)
        self._check(self.lib.sol_pack_phases(self.h, time, buf.ctypes.data, buf.size, C.byref(nb)))
        return buf.tobytes()

    def write_phases(self, path: str, time: float) -> None:
        self._check(self.lib.sol_write_phases(self.h, path.encode(), time))

    # ---- transfers ----
    def download(self, what: int) -> np.ndarray:
        n = self.n
        c = self.counts
        if what in (Y0, Y, ACCEL, YSCALE):
            out = np.empty((n, 6))
        elif what in (NN_INDEX, MIGTYPE, TYPE, ID):
            out = np.empty(n, dtype=np.int32)
        elif what == ACCEL_GASDRAG:
            out = np.zeros((int(c[4] + c[5]), 3))
        elif what == ACCEL_MIGTYPE1:
            out = np.zeros((int(c[2] + c[3]), 3))
        elif what == ACCEL_MIGTYPE2:
            out = np.zeros((int(c[1]), 3))
        else:
            out = np.empty(n)
        if out.size:
            self._check(self.lib.sol_download(self.h, what, out.ctypes.data))
        return out

    def upload(self, what: int, arr: np.ndarray):
        dt = np.int32 if what in (NN_INDEX, MIGTYPE, TYPE, ID) else np.float64
        arr = np.ascontiguousarray(arr, dtype=dt)
        self._check(self.lib.sol_upload(self.h, what, arr.ctypes.data))

    def flush_tiny(self, threshold: float = 1.0e-50):
        self._check(self.lib.sol_flush_tiny(self.h, threshold))

    # ---- multi-GPU ----
    @staticmethod
    def nccl_unique_id() -> bytes:
        buf = C.create_string_buffer(128)
        if load_library().sol_nccl_unique_id(buf) != 0:
            raise SolarisError(load_library().sol_last_error(None).decode())
        return buf.raw

    def dist_init(self, rank: int, nranks: int, unique_id: bytes):
        buf = C.create_string_buffer(unique_id, 128)
        self._check(self.lib.sol_dist_init(self.h, rank, nranks, buf))

    def shard_range(self):
        lo, hi = C.c_int(0), C.c_int(0)
        self._check(self.lib.sol_shard_range(self.h, C.byref(lo), C.byref(hi)))
        return lo.value, hi.value

    def gather_state(self):
        self._check(self.lib.sol_gather_state(self.h))

    # ---- measurement ----
    def time_gravity_kernel(self, reps: int):
        ms = C.c_float(0.0); pairs = C.c_double(0.0)
        self._check(self.lib.sol_time_gravity_kernel(self.h, reps, C.byref(ms), C.byref(pairs)))
        return ms.value, pairs.value

    def measure_fp64_peak(self) -> float:
        v = C.c_double(0.0)
        self._check(self.lib.sol_measure_fp64_peak(self.h, C.byref(v)))
        return v.value

    def selftest_fast_paths(self, samples: int, seed: int = 1) -> int:
        """Mismatches between the straight-line sqrt / reciprocal fast paths and the library's results (must be 0)."""
        bad = C.c_ulonglong(0)
        self._check(self.lib.sol_selftest_fast_paths(self.h, C.c_ulonglong(seed), C.c_longlong(samples), C.byref(bad)))
        return int(bad.value)

    def launch_count(self) -> int:
        return int(self.lib.sol_launch_count(self.h))

    def profile_enable(self, on: bool):
        self._check(self.lib.sol_profile_enable(self.h, int(on)))

    def profile_read(self, reset: bool = True):
        ms = (C.c_double * 6)(); n = (C.c_longlong * 6)()
        self._check(self.lib.sol_profile_read(self.h, ms, n, int(reset)))
        return list(ms), list(n)
