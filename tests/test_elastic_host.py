"""CPU tests of the elastic terms: the product's device header idp_b200/csrc/shell_elastic.cuh compiled
for the host (test-only harness, tests/host_shim/pair_host.cpp) against the oracle (oracle/orc_elastic.hpp: the reference's
energy definitions differentiated by second-order jets, cyclic-Jacobi makePD) and against finite differences."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


@pytest.fixture(scope="module")
def hs():
    src = os.path.join(ROOT, "tests", "host_shim", "pair_host.cpp")
    out = os.path.join(ROOT, "tests", "host_shim", "libpair_host.so")
    subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-Wno-unknown-pragmas", "-o", out, src])
    L = C.CDLL(out)
    L.hs_hinge_EgH.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    L.hs_membrane_EgH.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    return L


def _hinge(hs, x, thetabar, coef, proj):
    E = np.zeros(1); g = np.zeros(12); H = np.zeros((12, 12))
    assert hs.hs_hinge_EgH(_p(x), thetabar, coef, int(proj), _p(E), _p(g), _p(H)) == 0
    return E[0], g, H


def _membrane(hs, x, ib, coef, lam, mu, proj):
    E = np.zeros(1); g = np.zeros(9); H = np.zeros((9, 9))
    rc = hs.hs_membrane_EgH(_p(x), _p(ib), coef, lam, mu, int(proj), _p(E), _p(g), _p(H))
    return rc, E[0], g, H


def _rel(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


def _random_hinges(rng, n):
    out = []
    for _ in range(n):
        e0, e1 = rng.normal(size=3), rng.normal(size=3)
        ax = e1 - e0
        t = np.cross(ax, rng.normal(size=3)); t /= np.linalg.norm(t)
        ang = rng.uniform(-3.0, 3.0) if rng.random() < 0.8 else rng.choice([1e-7, -1e-6, 3.1, -3.1])
        k = ax / np.linalg.norm(ax)
        t2 = t * np.cos(ang) + np.cross(k, t) * np.sin(ang)
        m = 0.5 * (e0 + e1)
        x0 = m + rng.uniform(0.3, 2.0) * t + rng.uniform(-0.4, 0.4) * ax
        x3 = m - rng.uniform(0.3, 2.0) * t2 + rng.uniform(-0.4, 0.4) * ax
        out.append(np.concatenate([x0, e0, e1, x3]) * rng.choice([1e-2, 1.0, 30.0]))
    return out


def test_hinge_matches_oracle_jets(orc, hs):
    rng = np.random.default_rng(5)
    for x in _random_hinges(rng, 200):
        thetabar, coef = rng.uniform(-1, 1), 10.0 ** rng.uniform(-4, 3)
        info = np.array([[thetabar, 2.0, 0.5]])
        for proj in (False, True):
            E, g, H = _hinge(hs, x, thetabar, coef, proj)
            oE, og, oH, act = orc.hinge_batch(x.reshape(4, 3), np.arange(4), info, coef * 0.5 / 2.0, project_spd=proj)
            assert act[0] == 1
            assert abs(E - oE[0]) <= 1e-10 * abs(oE[0]) + 1e-300
            assert _rel(g, og.ravel()) <= 1e-10
            assert _rel(H, oH[0]) <= 1e-10, (proj, _rel(H, oH[0]))
            assert np.allclose(H, H.T, rtol=0, atol=1e-13 * np.abs(H).max())
            if proj:
                assert np.linalg.eigvalsh(H).min() >= -1e-12 * np.abs(H).max()
        # translation invariance and the sign convention of the angle
        assert np.abs(g.reshape(4, 3).sum(0)).max() <= 1e-12 * np.abs(g).max()


def test_hinge_angle_sign_and_finite_differences(orc, hs):
    x = np.array([0.0, 1, 0.3, 0, 0, 0, 1, 0, 0, 0.2, -1, 0.4])  # x0 ; x1, x2 ; x3
    th = orc.dihedral_angle(x)
    n1 = np.cross(x[3:6] - x[0:3], x[6:9] - x[0:3]); n2 = np.cross(x[6:9] - x[9:12], x[3:6] - x[9:12])
    c = n1 @ n2 / np.linalg.norm(n1) / np.linalg.norm(n2)
    assert abs(abs(th) - np.arccos(c)) < 1e-15 and th != 0
    E0, g, H = _hinge(hs, x, 0.1, 2.0, False)
    assert abs(E0 - 2.0 * (th - 0.1) ** 2) <= 1e-14
    eps = 1e-6
    for i in range(12):
        xp, xm = x.copy(), x.copy(); xp[i] += eps; xm[i] -= eps
        Ep, gp, _ = _hinge(hs, xp, 0.1, 2.0, False); Em, gm, _ = _hinge(hs, xm, 0.1, 2.0, False)
        assert abs((Ep - Em) / (2 * eps) - g[i]) <= 1e-7 * max(1.0, abs(g[i]))
        assert np.abs((gp - gm) / (2 * eps) - H[i]).max() <= 1e-6 * max(1.0, np.abs(H).max())


def test_membrane_matches_oracle_jets(orc, hs):
    rng = np.random.default_rng(7)
    for it in range(300):
        X0 = rng.normal(size=(3, 3)) * rng.choice([1e-2, 1.0, 10.0])
        e1, e2 = X0[1] - X0[0], X0[2] - X0[0]
        ib = np.array([e1 @ e1, e1 @ e2, e2 @ e2])
        x = X0 + rng.normal(size=(3, 3)) * 0.3 * np.sqrt(ib[0])   # stretched, sheared or compressed, never inverted in 2-D terms
        lam, mu, coef = 10.0 ** rng.uniform(-1, 4), 10.0 ** rng.uniform(-1, 4), 10.0 ** rng.uniform(-6, 0)
        for proj in (False, True):
            rc, E, g, H = _membrane(hs, x.ravel().copy(), ib, coef, lam, mu, proj)
            assert rc == 0
            oE, og, oH, act = orc.membrane_batch(x, np.arange(3), ib, coef, lam, mu, project_spd=proj)
            assert act[0] == 1
            assert abs(E - oE[0]) <= 1e-10 * max(abs(oE[0]), coef * mu * 1e-3)
            assert _rel(g, og.ravel()) <= 1e-10
            assert _rel(H, oH[0]) <= 1e-10, (it, proj, _rel(H, oH[0]))
            if proj:
                assert np.linalg.eigvalsh(H).min() >= -1e-12 * np.abs(H).max()
        assert np.abs(g.reshape(3, 3).sum(0)).max() <= 1e-12 * np.abs(g).max()
    # rest state: zero energy and gradient; degenerate rest triangle: skipped
    rc, E, g, H = _membrane(hs, X0.ravel().copy(), ib, 1.0, 3.0, 2.0, True)
    assert rc == 0 and abs(E) < 1e-13 and np.abs(g).max() < 1e-10 * np.sqrt(ib[0])
    rc, *_ = _membrane(hs, X0.ravel().copy(), np.array([1.0, 2.0, 4.0]), 1.0, 3.0, 2.0, True)
    assert rc == 2


def test_oracle_hinge_is_pinned_by_the_reference_dihedral_code(orc):
    """oracle/orc_elastic.hpp (jets of an atan2 formulation) against the REFERENCE's own Math/DIHEDRAL_ANGLE.h compiled in
    oracle/_ref/libidp_ref.so: angle bit for bit, gradient and Hessian of W = c (theta - thetabar)^2 built from the
    reference's closed forms (BENDING.h:196-200, 470-477) within 1e-10."""
    lib = os.path.join(ROOT, "oracle", "_ref", "libidp_ref.so")
    if not os.path.exists(lib):
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    L = C.CDLL(lib)
    if not hasattr(L, "ref_dihedral"):
        pytest.skip("oracle/_ref predates ref_dihedral")
    L.ref_dihedral.argtypes = [C.c_void_p] * 4
    rng = np.random.default_rng(17)
    for x in _random_hinges(rng, 300):
        th = np.zeros(1); g = np.zeros(12); H = np.zeros((12, 12))
        L.ref_dihedral(_p(x), _p(th), _p(g), _p(H))
        assert orc.dihedral_angle(x) == th[0]
        thetabar, c = rng.uniform(-1, 1), 10.0 ** rng.uniform(-3, 3)
        oE, og, oH, _ = orc.hinge_batch(x.reshape(4, 3), np.arange(4), np.array([[thetabar, 2.0, 0.5]]), c * 0.5 / 2.0, project_spd=False)
        d = th[0] - thetabar
        assert abs(oE[0] - c * d * d) <= 1e-12 * c * d * d
        assert _rel(og.ravel(), 2 * c * d * g) <= 1e-10
        assert _rel(oH[0], 2 * c * (d * H + np.outer(g, g))) <= 1e-10


def _ref_shell():
    lib = os.path.join(ROOT, "oracle", "_ref", "libidp_ref_shell.so")
    if not os.path.exists(lib):
        pytest.skip("oracle/_ref/libidp_ref_shell.so not built (needs /root/reference)")
    L = C.CDLL(lib)
    L.refshell_membrane.restype = C.c_long
    L.refshell_membrane.argtypes = [C.c_int, C.c_void_p, C.c_int] + [C.c_void_p] * 6 + [C.c_double, C.c_int, C.c_void_p, C.c_void_p, C.c_long] + [C.c_void_p] * 3
    L.refshell_hinges.restype = C.c_long
    L.refshell_hinges.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_void_p, C.c_double, C.c_int, C.c_void_p,
                                  C.c_void_p, C.c_long] + [C.c_void_p] * 3
    return L


def _blocks_to_csr(H, verts, n):
    import scipy.sparse as sp
    k = verts.shape[1]
    dof = (3 * verts[:, :, None] + np.arange(3)[None, None, :]).reshape(len(verts), 3 * k)
    r = np.repeat(dof[:, :, None], 3 * k, axis=2).ravel()
    c = np.repeat(dof[:, None, :], 3 * k, axis=1).ravel()
    return sp.coo_matrix((H.ravel(), (r, c)), shape=(n, n)).tocsr()


def test_oracle_elastic_terms_are_pinned_by_the_reference_shell_headers(orc):
    """oracle/orc_elastic.hpp against the REFERENCE's own FEM/Shell/MEMBRANE.h and BENDING.h (KL = false) compiled in
    oracle/_ref/libidp_ref_shell.so: Compute_Membrane_* / Compute_Bending_* energy, gradient and PSD-projected Hessian triplets on a
    deformed mesh with a Dirichlet mask (all-Dirichlet elements skipped by both), 1e-10."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from conftest import make_cases
    from shell_np import first_fundamental_forms, hinges
    L = _ref_shell()
    name, m, _d, _ = make_cases()[1]
    rng = np.random.default_rng(23)
    F = np.ascontiguousarray(m.btri[:, :3], np.int32)
    X0 = np.ascontiguousarray(m.X0)
    ib = np.ascontiguousarray(first_fundamental_forms(X0, F))
    X = np.ascontiguousarray(X0 + rng.normal(0, 0.03 * np.sqrt(ib[:, 0].mean()), X0.shape))
    dbc = (rng.uniform(size=m.nV) < 0.15).astype(np.uint8)
    area = 0.5 * np.linalg.norm(np.cross(X0[F[:, 1]] - X0[F[:, 0]], X0[F[:, 2]] - X0[F[:, 0]]), axis=1)
    vol = np.ascontiguousarray(area * 1e-2)
    lam = np.full(len(F), 1e4 * 0.4 / (1 - 0.16)); mu = np.full(len(F), 1e4 / 2.8)
    h, n3 = 0.04, 3 * m.nV
    for proj in (0, 1):
        E = C.c_double(0.0); g = np.zeros_like(X); cap = 81 * len(F)
        tr = np.zeros(cap, np.int32); tc = np.zeros(cap, np.int32); tv = np.zeros(cap)
        nt = L.refshell_membrane(m.nV, _p(X), len(F), _p(F), _p(ib), _p(vol), _p(lam), _p(mu), _p(dbc), h, proj, C.byref(E), _p(g), cap, _p(tr), _p(tc), _p(tv))
        oE, og, oH, act = orc.membrane_batch(X, F, ib, h * h * vol, lam, mu, dbc=dbc, project_spd=bool(proj))
        assert 0 < act.sum() < len(F) and nt == 81 * act.sum()
        assert abs(E.value - oE.sum()) <= 1e-10 * np.abs(oE).sum()
        assert np.abs(g - og).max() <= 1e-10 * np.abs(og).max()
        A = sp.coo_matrix((tv[:nt], (tr[:nt], tc[:nt])), shape=(n3, n3)).tocsr()
        B = _blocks_to_csr(oH, F, n3)
        assert spla.norm(A - B) <= 1e-10 * spla.norm(B), ("membrane", proj, spla.norm(A - B) / spla.norm(B))
    st, info = hinges(X0, F)
    info[:, 0] += rng.normal(0, 0.05, len(info))
    st = np.ascontiguousarray(st); info = np.ascontiguousarray(info)
    k = 1e4 * 1e-6 / (24 * (1 - 0.16))
    for proj in (0, 1):
        E = C.c_double(0.0); g = np.zeros_like(X); cap = 144 * len(st)
        tr = np.zeros(cap, np.int32); tc = np.zeros(cap, np.int32); tv = np.zeros(cap)
        nt = L.refshell_hinges(m.nV, _p(X), len(st), _p(st), _p(info), k, 1.0, _p(dbc), h, proj, C.byref(E), _p(g), cap, _p(tr), _p(tc), _p(tv))
        oE, og, oH, act = orc.hinge_batch(X, st, info, h * h * k, dbc=dbc, project_spd=bool(proj))
        assert nt == 144 * act.sum()
        assert abs(E.value - oE.sum()) <= 1e-10 * np.abs(oE).sum()
        assert np.abs(g - og).max() <= 1e-10 * np.abs(og).max()
        A = sp.coo_matrix((tv[:nt], (tr[:nt], tc[:nt])), shape=(n3, n3)).tocsr()
        B = _blocks_to_csr(oH, st, n3)
        assert spla.norm(A - B) <= 1e-10 * spla.norm(B), ("hinge", proj, spla.norm(A - B) / spla.norm(B))
