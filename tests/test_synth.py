import numpy as np

from solaris_b200 import synth


def test_sorted_by_body_type_and_counts():
    s = synth.mixed([1, 2, 3, 5, 4, 20, 10])
    assert s.n == 45 and list(s.type[:3]) == [1, 2, 2]
    assert np.all(np.diff(s.type) >= 0), "bodies must be sorted by BodyType (SURVEY.md Q2)"
    assert np.all(s.mass[s.type == synth.TEST] == 0) and np.all(s.gammaStokes[s.type == synth.TEST] == 0)
    assert s.y0.shape == (45, 6) and np.all(s.y0[0] == 0)


def test_deterministic_and_prefix_reproducible():
    a = synth.massive_disk(1000)
    b = synth.massive_disk(1000)
    assert np.array_equal(a.y0, b.y0) and np.array_equal(a.mass, b.mass)
    c = synth.massive_disk(400)
    assert np.array_equal(a.y0[:400], c.y0) and np.array_equal(a.mass[:400], c.mass)


def test_bound_orbits_and_barycentre():
    s = synth.trojans(500)
    r = np.sqrt((s.y0[1:, :3] ** 2).sum(1)); v2 = (s.y0[1:, 3:] ** 2).sum(1)
    assert np.all(0.5 * v2 - synth.GAUSS2 * (1 + s.mass[1:]) / r < 0)
    b = synth.to_barycentric(synth.solar_system())
    M = 9
    assert np.abs((b.mass[:M, None] * b.y0[:M]).sum(0)).max() < 1e-18
    assert np.array_equal(b["y0"], b.y0), "attribute and key access must alias"


def test_pairs_per_eval_matches_reference_loops():
    c = [1, 2, 3, 5, 4, 20, 10]
    n, M, s = 45, 11, 4
    ac = sum((M + s - 2) if i < M else (M - 1) for i in range(1, n))
    assert synth.pairs_per_eval(c, False) == ac
    assert synth.pairs_per_eval(c, True) == n * M - M
