"""Known-answer vectors held by the reference's OWN tests for functions on the hot path
(SURVEY.md §8c items 2 and 4), applied to the oracle's restatement of those functions:
  * src/Solaris.NBody.Cuda.Test/unit_test.cpp:558-699  circular_velocity, gas_velocity, gas_density_at
  * Test/Test.cpp:421-532                              mean free path, temperature, mean thermal speed
"""
import ctypes as C

import numpy as np

import oraclelib
from oraclelib import NebulaPod, default_nebula


def _lib():
    oraclelib.ensure_oracle_built()
    L = C.CDLL(oraclelib.ORACLE_SO)
    P = C.POINTER(NebulaPod)
    L.oracle_circular_velocity.argtypes = [C.c_double] * 3 + [C.POINTER(C.c_double)]
    L.oracle_gas_velocity.argtypes = [P] + [C.c_double] * 3 + [C.POINTER(C.c_double)]
    L.oracle_gas_density_at.argtypes = [P] + [C.c_double] * 3
    L.oracle_gas_density_at.restype = C.c_double
    for f in (L.oracle_temperature_cmu, L.oracle_mean_thermal_speed_cmu):
        f.argtypes = [P, C.c_double, C.c_double]
        f.restype = C.c_double
    L.oracle_mean_free_path.argtypes = [P, C.c_double]
    L.oracle_mean_free_path.restype = C.c_double
    L.oracle_reduction_factor.argtypes = [P, C.c_double]
    L.oracle_reduction_factor.restype = C.c_double
    L.oracle_orbital_element_ae.argtypes = [C.c_double] + [C.POINTER(C.c_double)] * 3
    return L


def test_circular_velocity_known_answers():
    L = _lib()
    out = (C.c_double * 2)()
    for (x, y), exp in (((1.0, 0.0), (0.0, 1.0)), ((0.0, 1.0), (-1.0, 0.0)), ((-1.0, 0.0), (0.0, -1.0)), ((0.0, -1.0), (1.0, 0.0))):
        L.oracle_circular_velocity(1.0, x, y, out)          # unit_test.cpp:564-594: exact
        assert (out[0], out[1]) == exp
    L.oracle_circular_velocity(0.001, 5.0, 0.5, out)         # unit_test.cpp:596-607: 1e-15
    assert abs(out[0] - (-0.0014036989255830)) <= 1e-15 and abs(out[1] - 0.01403698925583099) <= 1e-15


def test_gas_velocity_known_answer():
    L = _lib()
    p = default_nebula()
    p.eta_c, p.eta_index = 1.0e-3, 0.5
    out = (C.c_double * 2)()
    L.oracle_gas_velocity(C.byref(p), 1.0, 1.0, 0.0, out)    # unit_test.cpp:627-634: exact
    assert out[0] == 0.0 and out[1] == 0.99899949949937412368543414284205


def test_gas_density_known_answers():
    L = _lib()
    p = default_nebula()
    p.density_c, p.density_index = 1.0e-10, -3.0
    p.scale_height_c, p.scale_height_index = 5.0e-2, 1.5
    p.inner_edge = 0.1        # the CUDA-side tests use a 0.1 AU inner edge
    f = lambda x, y, z: L.oracle_gas_density_at(C.byref(p), x, y, z)   # noqa: E731
    assert abs(f(0.1, 0.0, 0.0) - 1.0e-7) <= 1e-15           # unit_test.cpp:659-665
    assert abs(f(0.05, 0.0, 0.0) - 6.25e-9) <= 1e-15          # :667-673
    assert f(1.0, 0.0, 0.0) == 1.0e-10                        # :675-681
    assert f(0.0, 1.0, 0.0) == 1.0e-10                        # :683-689
    assert abs(f(1.0, 0.0, 5.0e-2) - 3.6787944117144232159552377016146e-11) <= 1e-16   # :691-697


def test_temperature_thermal_speed_mean_free_path():
    L = _lib()
    p = default_nebula()
    assert abs(L.oracle_temperature_cmu(C.byref(p), 1.0, 1.0) - 98.903471085889933) <= 1.0      # Test.cpp:489-497
    assert abs(L.oracle_temperature_cmu(C.byref(p), 1.0, 0.1) - 312.67170117983335) <= 1.0      # Test.cpp:499-510
    meter_to_au, second_to_day = 1.0 / 1.495978707e11, 1.0 / 86400.0
    exp = 950.7244052398592 * meter_to_au / second_to_day                                      # Test.cpp:524-530
    assert abs(L.oracle_mean_thermal_speed_cmu(C.byref(p), 1.0, 1.0) - exp) <= 1e-5
    # Test.cpp:436-449: the ctor-time mean free path law at 1 AU equals MeanFreePath_CMU(rho_c)
    proton_cmu = 1.672621777e-27 / 1.98911e30
    exp_l = 2.3 * proton_cmu / (np.sqrt(2.0) * np.pi * (3.0e-10 * meter_to_au) ** 2 * p.density_c)
    assert abs(L.oracle_mean_free_path(C.byref(p), 1.0) - exp_l) <= 1e-5 * exp_l


def test_reduction_factor_branches():
    L = _lib()
    p = default_nebula()
    assert L.oracle_reduction_factor(C.byref(p), 123.0) == 1.0            # CONSTANT
    p.decrease_type, p.t0, p.t1 = 1, 10.0, 20.0                           # LINEAR, GasComponent.cpp:43-53
    assert L.oracle_reduction_factor(C.byref(p), 5.0) == 1.0
    assert L.oracle_reduction_factor(C.byref(p), 15.0) == 0.5
    assert L.oracle_reduction_factor(C.byref(p), 25.0) == 0.0
    p.decrease_type, p.time_scale = 2, 100.0                              # EXPONENTIAL
    assert L.oracle_reduction_factor(C.byref(p), 100.0) == np.exp(-1.0)


def test_orbital_elements_circular_orbit():
    """Ephemeris::CalculateOrbitalElement on a circular orbit: a = r, e = 0 via the |e2| < 1e-14 clamp
    that only exists with the MSVC abs() semantics."""
    L = _lib()
    mu = 2.959122082855911025e-4
    rv = (C.c_double * 6)(2.0, 0.0, 0.0, 0.0, np.sqrt(mu / 2.0), 0.0)
    a, e = C.c_double(0), C.c_double(0)
    assert L.oracle_orbital_element_ae(mu, rv, C.byref(a), C.byref(e)) == 0
    assert abs(a.value - 2.0) <= 1e-14 and e.value == 0.0
    rv[4] = 2.0 * np.sqrt(mu / 2.0)       # hyperbolic: energy >= 0 -> returns 1, a and e untouched
    a.value = e.value = -7.0
    assert L.oracle_orbital_element_ae(mu, rv, C.byref(a), C.byref(e)) == 1
    assert a.value == -7.0 and e.value == -7.0
