"""Lagged friction: the product's device header idp_b200/csrc/friction.cuh compiled for the host (CPU
test) and the CUDA path through the C ABI (GPU test) against the REFERENCE's own FEM/FRICTION.h compiled in
oracle/_ref/libidp_ref_ipc.so (Compute_Friction_Basis / _Potential / _Gradient / _Hessian) -- or, where that build is absent,
against the oracle restatement oracle/orc_friction.hpp, which is itself pinned by the reference build here. Bars: friction rows
identical, closest points / bases / normal forces / E / g / H within 1e-10 relative."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from conftest import make_cases  # noqa: E402
from oracle import ref_binding  # noqa: E402

KAPPA, MU, EPSV2H2 = 1e5, 0.4, 1e-4 * 0.01 ** 2 * 25.0
HAVE_REF = ref_binding.ipc_available() and hasattr(C.CDLL(ref_binding.LIB_IPC), "refipc_friction")
needs_ref = pytest.mark.skipif(not HAVE_REF, reason="oracle/_ref with the reference's FRICTION.h is not built (needs /root/reference)")


def _checker(orc):
    """the reference's own friction code where it is built, else its restatement"""
    return ref_binding.ReferenceIPC() if HAVE_REF else orc


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _scene(orc, case, k):
    """contact rows of a test mesh (oracle), a step-start state Xn and an iterate X that slides tangentially"""
    name, m, d, dhats = case
    dh2 = dhats[-1] ** 2
    om = orc.mesh(m.X, m.X0, m.bnode, m.bedge, m.btri, m.dbc)
    rows, info, _, _ = orc.constraint_set(om, dh2)
    rng = np.random.default_rng(100 + k)
    scale = np.sqrt(EPSV2H2)
    Xn = m.X - rng.normal(0, 3.0 * scale, m.X.shape)           # displacements on both sides of the eps_v h clamp
    X = m.X + rng.normal(0, 0.3 * scale, m.X.shape)
    some = rng.uniform(size=len(X)) < 0.2
    X[some] = Xn[some]                                          # vertices that did not move at all
    return m, rows, dh2, Xn, X


@needs_ref
def test_oracle_friction_matches_reference(orc):
    """oracle/orc_friction.hpp against the reference's FRICTION.h: rows and normal forces bit for bit, the rest 1e-12."""
    ref = ref_binding.ReferenceIPC()
    for k, case in enumerate(make_cases()):
        m, rows, dh2, Xn, X = _scene(orc, case, k)
        r = ref.friction(m.X, rows, dh2, KAPPA, X=X, Xn=Xn, epsv2_h2=EPSV2H2, mu=MU)
        o = orc.friction(m.X, rows, dh2, KAPPA, X=X, Xn=Xn, epsv2_h2=EPSV2H2, mu=MU)
        assert np.array_equal(o["rows"], r["rows"]) and np.array_equal(o["normal_force"], r["normal_force"])
        four = (r["rows"][:, 0] >= 0) | (r["rows"][:, 3] >= 0)                         # EE / PT: two parameters
        pe = (r["rows"][:, 0] < 0) & (r["rows"][:, 2] >= 0) & (r["rows"][:, 3] < 0)        # PE: one (the reference leaves the other unset)
        assert np.array_equal(o["closest"][four], r["closest"][four]) and np.array_equal(o["closest"][pe, 0], r["closest"][pe, 0])
        assert np.allclose(o["basis"], r["basis"], rtol=0, atol=1e-15)
        assert abs(o["E"] - r["E"]) <= 1e-13 * abs(r["E"]) and np.abs(o["g"] - r["g"]).max() <= 1e-12 * np.abs(r["g"]).max()
        n3 = 3 * len(X)
        A = sp.coo_matrix((o["triplets"][2], (o["triplets"][0], o["triplets"][1])), shape=(n3, n3)).tocsr()
        B = sp.coo_matrix((r["triplets"][2], (r["triplets"][0], r["triplets"][1])), shape=(n3, n3)).tocsr()
        assert len(o["triplets"][2]) == len(r["triplets"][2]) and spla.norm(A - B) <= 1e-12 * spla.norm(B)


def test_friction_header_on_host_matches_reference(orc):
    src = os.path.join(ROOT, "tests", "host_shim", "pair_host.cpp")
    out = os.path.join(ROOT, "tests", "host_shim", "libpair_host.so")
    subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-Wno-unknown-pragmas", "-o", out, src])
    hs = C.CDLL(out)
    hs.hs_friction.argtypes = [C.c_int] + [C.c_void_p] * 4 + [C.c_double] * 5 + [C.c_void_p] * 9
    ref = _checker(orc)
    for k, case in enumerate(make_cases()):
        m, rows, dh2, Xn, X = _scene(orc, case, k)
        n = len(rows)
        assert n > 0
        r = ref.friction(m.X, rows, dh2, KAPPA, X=X, Xn=Xn, epsv2_h2=EPSV2H2, mu=MU)
        nv = np.zeros(n, np.int32); verts = np.zeros((n, 4), np.int32); w = np.zeros((n, 4)); basis = np.zeros((n, 6)); lam = np.zeros(n)
        cp = np.zeros((n, 2)); E = np.zeros(1); g = np.zeros_like(X); H = np.zeros((n, 12, 12))
        rows_c = np.ascontiguousarray(rows, np.int32); Xb = np.ascontiguousarray(m.X)
        hs.hs_friction(n, _p(rows_c), _p(Xb), _p(X), _p(Xn), dh2, KAPPA, 0.0, np.sqrt(EPSV2H2), MU, _p(nv), _p(verts), _p(w), _p(basis), _p(lam), _p(cp),
                       _p(E), _p(g), _p(H))
        act = nv > 0
        assert act.sum() == len(r["rows"]) and np.array_equal(rows_c[act], r["rows"])       # the non-mollified rows, in order
        mult = np.where(rows_c[act][:, 3] < -1, -rows_c[act][:, 3], 1)
        assert np.allclose(lam[act] / mult, r["normal_force"], rtol=1e-12, atol=0)
        assert np.allclose(basis[act], r["basis"], rtol=0, atol=1e-12)
        four = nv[act] == 4
        assert np.allclose(cp[act][four], r["closest"][four], rtol=1e-10, atol=1e-12)
        assert abs(E[0] - r["E"]) <= 1e-10 * abs(r["E"])
        assert np.abs(g - r["g"]).max() <= 1e-10 * np.abs(r["g"]).max()
        dof = (3 * verts[:, :, None] + np.arange(3)[None, None, :]).reshape(n, 12)
        rr = np.repeat(dof[:, :, None], 12, axis=2); cc = np.repeat(dof[:, None, :], 12, axis=1)
        keep = np.abs(H) > 0
        A = sp.coo_matrix((H[keep], (rr[keep], cc[keep])), shape=(3 * len(X),) * 2).tocsr()
        tr, tc, tv = r["triplets"]
        B = sp.coo_matrix((tv, (tr, tc)), shape=A.shape).tocsr()
        assert spla.norm(A - B) <= 1e-10 * spla.norm(B), (case[0], spla.norm(A - B) / spla.norm(B))
        assert spla.norm(B) > 0


@pytest.mark.gpu
def test_friction_on_the_device_matches_reference(lib_built, orc):
    from idp_b200 import ContactContext
    ref = _checker(orc)
    for k, case in enumerate(make_cases()):
        m, rows, dh2, Xn, X = _scene(orc, case, k)
        n3 = 3 * m.nV
        c = ContactContext(0)
        try:
            c.set_surface_mesh(m)
            assert c.constraint_set(dh2) == len(rows)
            grows, _ = c.get_constraints()
            nfr = c.friction_update(dh2, KAPPA)                      # Compute_Friction_Basis at the current positions
            r = ref.friction(m.X, grows, dh2, KAPPA, X=X, Xn=Xn, epsv2_h2=EPSV2H2, mu=MU)
            assert nfr == len(r["rows"])
            frows, cp, basis, nf = c.get_friction()
            assert np.array_equal(frows, r["rows"])
            assert np.allclose(nf, r["normal_force"], rtol=1e-12, atol=0) and np.allclose(basis, r["basis"], rtol=0, atol=1e-12)
            c.friction_set(Xn, EPSV2H2, MU)
            c.set_positions(X)
            E = c.friction_energy(E0=0.5)
            assert abs(E - 0.5 - r["E"]) <= 1e-10 * abs(r["E"])
            g = c.friction_gradient()
            assert np.abs(g - r["g"]).max() <= 1e-10 * np.abs(r["g"]).max()
            # the friction blocks ride along with the barrier Hessian: H(mu) - H(0) = the reference's friction triplets
            c.set_constraints(grows)
            p1, c1, v1 = c.barrier_hessian(dh2, KAPPA, project_spd=True)
            c.friction_set(None, 0.0, 0.0)
            p0, c0, v0 = c.barrier_hessian(dh2, KAPPA, project_spd=True)
            A = sp.csr_matrix((v1, c1, p1), shape=(n3, n3)) - sp.csr_matrix((v0, c0, p0), shape=(n3, n3))
            tr, tc, tv = r["triplets"]
            B = sp.coo_matrix((tv, (tr, tc)), shape=(n3, n3)).tocsr()
            assert spla.norm(A - B) <= 1e-9 * spla.norm(B), (case[0], spla.norm(A - B) / spla.norm(B))
            assert c.friction_energy() == 0.0                        # mu = 0 switches the term off
        finally:
            c.close()


@needs_ref
def test_friction_components_oracle_matches_reference(orc):
    """Compute_Friction_Coef: oracle/orc_friction.hpp against the reference's own function."""
    ref = ref_binding.ReferenceIPC()
    L, O = ref.lib, orc.lib
    L.refipc_friction_coef.restype = C.c_double
    L.refipc_friction_coef.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    O.orc_friction_coef.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    m, rows, dh2, Xn, X = _scene(orc, make_cases()[0], 0)
    r = ref.friction(m.X, rows, dh2, KAPPA)
    rng_ = np.array([m.nV // 3, 2 * m.nV // 3, m.nV], np.int32)
    muc = np.array([0.1, 0.2, 0.3, 0.2, 0.4, 0.5, 0.3, 0.5, 0.6])
    a, b = r["normal_force"].copy(), r["normal_force"].copy()
    fr = np.ascontiguousarray(r["rows"], np.int32)
    assert L.refipc_friction_coef(len(fr), _p(fr), 3, _p(rng_), _p(muc), _p(a)) == 1.0
    assert O.orc_friction_coef(len(fr), _p(fr), 3, _p(rng_), _p(muc), _p(b)) == 0
    assert np.array_equal(a, b) and not np.array_equal(a, r["normal_force"])


@pytest.mark.gpu
def test_friction_components_on_the_device(lib_built, orc):
    """idp_friction_set_components: the frozen normal forces carry the per-component coefficient (checked against the reference's
    Compute_Friction_Coef where built, else the oracle's)."""
    from idp_b200 import ContactContext
    m, rows, dh2, Xn, X = _scene(orc, make_cases()[0], 0)
    rng_ = np.array([m.nV // 3, 2 * m.nV // 3, m.nV], np.int32)
    muc = np.array([0.1, 0.2, 0.3, 0.2, 0.4, 0.5, 0.3, 0.5, 0.6])
    c = ContactContext(0)
    try:
        c.set_surface_mesh(m)
        c.constraint_set(dh2)
        c.friction_update(dh2, KAPPA)
        frows, _, _, nf0 = c.get_friction()
        c.friction_set_components(rng_, muc)
        c.friction_update(dh2, KAPPA)
        frows1, _, _, nf1 = c.get_friction()
        assert np.array_equal(frows, frows1)
        want = nf0.copy()
        fr = np.ascontiguousarray(frows, np.int32)
        if HAVE_REF:
            L = ref_binding.ReferenceIPC().lib
            L.refipc_friction_coef.restype = C.c_double
            L.refipc_friction_coef.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
            L.refipc_friction_coef(len(fr), _p(fr), 3, _p(rng_), _p(muc), _p(want))
        else:
            orc.lib.orc_friction_coef.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
            assert orc.lib.orc_friction_coef(len(fr), _p(fr), 3, _p(rng_), _p(muc), _p(want)) == 0
        assert np.allclose(nf1, want, rtol=1e-14, atol=0) and not np.allclose(nf1, nf0)
        c.friction_set_components(None, None)
        c.friction_update(dh2, KAPPA)
        assert np.array_equal(c.get_friction()[3], nf0)
        with pytest.raises(Exception):  # a vertex beyond the last bound: the reference prints "can't find node compI" and exits
            c.friction_set_components(np.array([m.nV // 2], np.int32), np.array([0.3]))
            c.friction_update(dh2, KAPPA)
    finally:
        c.close()
