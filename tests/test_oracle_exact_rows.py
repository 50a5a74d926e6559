"""The extended-precision row oracle (oracle/oracle.c: oracle_gravity_rows_exact, long double + Neumaier summation)
that the full-size GPU test measures the north star's 1e-13 against.  Pinned here on the CPU:
  * against IEEE binary128 (oracle_gravity_row_quad, libquadmath) - the two must round to the same double up to 1 ulp;
  * against the reference's own arithmetic (the bit-exact row restatement of GravityAC / GravityBC) on systems where the
    double-precision sum is well conditioned - there all three agree to a few 1e-16."""
import numpy as np
import pytest

from solaris_b200 import synth
from oraclelib import Oracle


@pytest.mark.parametrize("bary", [False, True])
def test_exact_rows_match_binary128_and_reference(bary):
    s = synth.mixed([1, 2, 3, 200, 40, 100, 54], seed=5)
    if bary:
        s = synth.to_barycentric(s)
    o = Oracle(s, bary, None)
    rows = np.arange(0 if bary else 1, s.n, dtype=np.int32)
    ex = o.gravity_rows_exact(s.y0, rows, threads=4)
    ref = o.gravity_rows(s.y0, 0, s.n, 1)[rows, 3:]
    nrm = np.sqrt((ex ** 2).sum(axis=1))
    assert np.all(nrm > 0)
    assert (np.abs(ref - ex).max(axis=1) / nrm).max() <= 2e-14      # 400 well-conditioned terms in double
    for k in list(range(0, len(rows), 37)) + [len(rows) - 1]:
        q = o.gravity_row_quad(s.y0, int(rows[k]))
        assert np.abs(q - ex[k]).max() <= 2.3e-16 * nrm[k]          # same double up to the final rounding


def test_exact_rows_of_a_large_disk_against_binary128():
    """N = 50 000 self-gravitating bodies: 5e4-term sums, where compensated long double and binary128 must still round
    to the same double, while the reference's sequential double sum is already ~1e-14 away."""
    s = synth.massive_disk(50_000)
    o = Oracle(s, False, None)
    rows = np.array([1, 777, 25_000, s.n - 1], dtype=np.int32)
    ex = o.gravity_rows_exact(s.y0, rows, threads=4)
    for k, i in enumerate(rows):
        q = o.gravity_row_quad(s.y0, int(i))
        nrm = np.sqrt((q ** 2).sum())
        assert np.abs(q - ex[k]).max() <= 2.3e-16 * nrm
        ref = o.gravity_rows(s.y0, int(i), int(i) + 1, 1)[0, 3:]
        assert np.abs(ref - ex[k]).max() <= 1e-13 * nrm


def test_astrocentric_star_row_is_zero():
    s = synth.massive_disk(300)
    o = Oracle(s, False, None)
    assert np.all(o.gravity_rows_exact(s.y0, np.array([0], dtype=np.int32)) == 0.0)
    assert np.all(o.gravity_row_quad(s.y0, 0) == 0.0)
