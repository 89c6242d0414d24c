"""The oracle against golden vectors produced by the REFERENCE's own per-pair code (tests/golden/ref_pair_math.npz,
made by tests/golden/make_golden.py from oracle/_ref = /root/reference/Library/Math compiled against stub Eigen
headers), and live against oracle/_ref when it is present. Integer results (distance types, AABB predicates, hit flags)
and exactly-specified arithmetic (distances, ACCD steps, barrier scalars) must match bit for bit; the derivative code
(MATLAB-generated in the reference, hand-derived here) to 1e-10 relative."""
import os

import numpy as np
import pytest

from conftest import ROOT

G = np.load(os.path.join(ROOT, "tests", "golden", "ref_pair_math.npz"))
NP = {0: 2, 1: 3, 2: 4, 3: 4, 6: 4}


def relerr(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


@pytest.mark.parametrize("kind", sorted(NP))
def test_distances_and_derivatives(orc, kind):
    for x, d, g, H in zip(G["x_%d" % kind], G["d_%d" % kind], G["g_%d" % kind], G["H_%d" % kind]):
        assert orc.dist2(kind, x) == d  # same expression, same rounding
        og, oH = orc.grad_hess(kind, x)
        assert relerr(og, g) <= 1e-10 and relerr(oH, H) <= 1e-10
        jg, jH = orc.grad_hess(kind, x, jet=True) if kind else (og, oH)
        assert relerr(jg, g) <= 1e-10 and relerr(jH, H) <= 1e-10


def test_distance_types_bit_exact(orc):
    for x, pt, ee, dpt, dee in zip(G["type_x"], G["pt_type"], G["ee_type"], G["pt_unclassified"], G["ee_unclassified"]):
        assert orc.pt_type(x) == pt and orc.ee_type(x) == ee
        a, b = orc.dist2(4, x), orc.dist2(5, x)
        assert (a == dpt or (np.isnan(a) and np.isnan(dpt))) and (b == dee or (np.isnan(b) and np.isnan(dee)))
    assert len(set(G["pt_type"])) == 7 and len(set(G["ee_type"])) >= 8  # the fixtures exercise every branch


def test_mollifier(orc):
    for x, eps, e, g, H, thr in zip(G["moll_x"], G["moll_eps"], G["moll_e"], G["moll_g"], G["moll_H"], G["moll_thr"]):
        oe, og, oH = orc.mollifier(x, eps)
        assert oe == e and orc.mollifier_threshold(x) == thr
        assert np.abs(og - g).max() <= 1e-10 * max(np.abs(g).max(), 1e-300)
        assert np.abs(oH - H).max() <= 1e-10 * max(np.abs(H).max(), 1e-300)


def test_barrier_scalars_bit_exact(orc):
    for d, ref in zip(G["bar_d"], G["bar"]):
        assert orc.barrier_scalar(d, 1e-4, 1e5) == tuple(ref)


def test_aabb_predicates_bit_exact(orc):
    for x, dd, dist, ref in zip(G["aabb_x"], G["aabb_d"], G["aabb_dist"], G["aabb"]):
        assert [orc.aabb(k, x, dd, dist) for k in range(4)] == list(ref.astype(bool))
    # ties pass: the rejection is a strict > (CCD.h:159)
    x = np.array([[0, 0, 1.5], [0, 0, 0], [1, 0, 0], [0, 1, 0]], np.float64)
    assert orc.aabb(0, x, None, 1.5) and not orc.aabb(0, x, None, np.nextafter(1.5, 0))


@pytest.mark.parametrize("name,kind", [("pt", 0), ("ee", 1)])
def test_additive_ccd_bit_exact(orc, name, kind):
    X, D, hit, toc = (G["accd_%s_%s" % (name, k)] for k in ("x", "d", "hit", "toc"))
    assert hit.sum() > 10 and (~hit).sum() > 10
    for i, (x, d, h, t) in enumerate(zip(X, D, hit, toc)):
        th = 1e-3 if i % 3 == 0 else 0.0
        oh, ot, _ = orc.accd(kind, x, d, 1.0, 0.1, th)
        assert oh == h
        if h:
            assert ot == t  # identical sequence of IEEE operations


def test_live_against_reference_build(orc):
    """When oracle/_ref exists (authoring container / shipped .so), re-check on fresh random inputs."""
    from oracle import ref_binding
    if not ref_binding.available():
        pytest.skip("oracle/_ref not built")
    ref = ref_binding.Reference()
    rng = np.random.default_rng(77)
    for _ in range(3000):
        x = rng.normal(size=(4, 3)) * 10.0 ** rng.uniform(-3, 1)
        assert orc.pt_type(x) == ref.pt_type(x) and orc.ee_type(x) == ref.ee_type(x)
        for k in (0, 1, 2, 3, 4, 5, 6):
            assert orc.dist2(k, x[: NP.get(k, 4)]) == ref.dist2(k, x[: NP.get(k, 4)])
    for _ in range(200):
        x = rng.normal(size=(4, 3)); d = rng.normal(size=(4, 3))
        for kind in (0, 1):
            h1, t1, _ = orc.accd(kind, x, d, 0.9)
            h2, t2 = ref.accd(kind, x, d, 0.9)
            assert h1 == h2 and (not h1 or t1 == t2)
        A = rng.normal(size=(12, 12)); A = A + A.T
        assert np.linalg.norm(orc.make_pd(A) - ref.make_pd12(A)) <= 1e-12 * np.linalg.norm(A)
