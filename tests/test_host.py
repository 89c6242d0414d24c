"""CPU tests of the host side: the C ABI library loads and exports every symbol the header declares (no compute call
without a GPU), the product fails loudly without a device, the device math headers agree with the oracle when compiled
for the host, and the input-ordering contract of the synthetic meshes."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import ROOT, make_cases


def test_library_exports_every_declared_symbol(lib_built):
    import idp_b200
    hdr = open(os.path.join(ROOT, "include", "idp_contact.h")).read()
    declared = sorted(set(re.findall(r"\b(idp_[a-z0-9_]+)\s*\(", hdr)))
    assert declared == sorted(idp_b200.EXPORTS)
    lib = C.CDLL(lib_built)
    for name in declared:
        assert getattr(lib, name) is not None, name
    # no torch types / C++ names in the ABI: every exported idp_ symbol is unmangled
    out = subprocess.run(["nm", "-D", "--defined-only", lib_built], capture_output=True, text=True).stdout
    exported = {ln.split()[-1] for ln in out.splitlines() if " T " in ln and "idp_" in ln.split()[-1][:4]}
    assert set(declared) <= exported


def test_no_cpu_fallback_without_device(lib_built):
    """On a machine without CUDA the product path must fail loudly, not compute on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from idp_b200 import ContactContext, IdpError
    with pytest.raises(IdpError) as e:
        ContactContext(0)
    assert e.value.code == 1  # IDP_ERR_CUDA


def test_product_never_touches_the_oracle():
    for base, _dirs, files in os.walk(os.path.join(ROOT, "idp_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", "Makefile")):
                txt = open(os.path.join(base, f), errors="ignore").read()
                assert "oracle" not in txt.replace("CPU oracle", "").replace("the oracle", "").replace("oracle:", "").lower() \
                    or f in ("psd_lowrank.cuh",) or "import oracle" not in txt, (base, f)
                assert "from oracle" not in txt and "import oracle" not in txt and "oracle/" not in txt.replace("oracle/_ref", ""), (base, f)


@pytest.fixture(scope="module")
def host_shim():
    src = os.path.join(ROOT, "tests", "host_shim", "pair_host.cpp")
    out = os.path.join(ROOT, "tests", "host_shim", "libpair_host.so")
    subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-Wno-unknown-pragmas",
                           "-o", out, src])
    hs = C.CDLL(out)
    hs.hs_dist2_unclassified.restype = C.c_double
    hs.hs_row_EgH.argtypes = [C.c_void_p] * 3 + [C.c_double] * 4 + [C.c_int, C.c_int] + [C.c_void_p] * 5
    hs.hs_accd.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_void_p, C.c_void_p]
    hs.hs_make_pd.argtypes = [C.c_int, C.c_void_p, C.c_void_p]
    return hs


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def test_device_classification_is_bit_exact_on_host(orc, host_shim):
    rng = np.random.default_rng(3)
    for t in range(20000):
        x = rng.normal(size=(4, 3))
        if t % 3 == 0:
            x[0] = x[1] + 1e-3 * rng.normal(size=3)
        if t % 5 == 0:
            x[3] = x[2] + (x[1] - x[0]) * (1 + 1e-9 * rng.normal())
        assert orc.pt_type(x) == host_shim.hs_pt_type(_p(x))
        assert orc.ee_type(x) == host_shim.hs_ee_type(_p(x))
        assert orc.dist2(4, x) == host_shim.hs_dist2_unclassified(0, _p(x))
        assert orc.dist2(5, x) == host_shim.hs_dist2_unclassified(1, _p(x))


def test_device_accd_is_bit_exact_on_host(orc, host_shim):
    rng = np.random.default_rng(4)
    for kind in (0, 1):
        hits = 0
        for _ in range(600):
            x = rng.normal(size=(4, 3)); d = rng.normal(size=(4, 3)) * rng.uniform(0.1, 3)
            hit, toc, its = orc.accd(kind, x, d, 1.0)
            t2 = C.c_double(0); it2 = C.c_int(0)
            r = host_shim.hs_accd(kind, _p(x), _p(d), 0.1, 0.0, 1.0, C.byref(t2), C.byref(it2))
            assert bool(r == 1) == hit
            if hit:
                hits += 1
                assert t2.value == toc and it2.value == its
        assert hits > 20


@pytest.mark.parametrize("lowrank", [0, 1])
def test_device_row_math_matches_oracle_on_host(orc, host_shim, lowrank):
    """pair_deriv.cuh (dense) and psd_lowrank.cuh (reduced coordinates + Jacobi) against the oracle, 1e-10 relative."""
    for name, m, d, dhats in make_cases()[:2]:
        om = orc.mesh(m.X, m.X0, m.bnode, m.bedge, m.btri, m.dbc)
        dh = dhats[-1]
        rows, info, _, _ = orc.constraint_set(om, dh * dh)
        rng = np.random.default_rng(1)
        seen = set()
        for i in rng.choice(len(rows), min(len(rows), 1200), replace=False):
            r = np.ascontiguousarray(rows[i])
            seen.add(tuple(int(v >= 0) for v in r))
            for spd in (0, 1):
                st, E, g, H, verts = orc.row_EgH(om, r, 1.0, dh * dh, 1e5, project_spd=bool(spd))
                E2 = C.c_double(0); g2 = np.zeros(12); H2 = np.zeros(144); nv = C.c_int(0); vv = np.zeros(4, np.int32)
                st2 = host_shim.hs_row_EgH(_p(r), _p(m.X), _p(m.X0), 1.0, dh * dh, 1e5, 0.0, spd, lowrank, C.byref(E2),
                                           _p(g2), _p(H2), C.byref(nv), _p(vv))
                n = 3 * nv.value
                assert st == st2 == 0 and np.array_equal(vv[:nv.value], verts)
                assert abs(E2.value - E) <= 1e-12 * abs(E)
                assert np.abs(g2[:n] - g).max() <= 1e-10 * np.abs(g).max()
                assert np.linalg.norm(H2[:n * n].reshape(n, n) - H) <= 1e-10 * np.linalg.norm(H)
        assert len(seen) >= 4, seen


@pytest.mark.parametrize("n", [6, 9])
def test_device_psd_projection_on_hard_spectra(host_shim, n):
    """make_pd_ql (psd_lowrank.cuh: Householder + pipelined implicit QL, compiled for the host) against numpy's eigh
    projection on spectra the mesh tests do not guarantee: repeated and clustered eigenvalues, exact zeros, rank one,
    already block-diagonal / diagonal input, 1e16 dynamic range, definite matrices, tiny and huge scales."""
    rng = np.random.default_rng(12)
    P = lambda a: a.ctypes.data_as(C.c_void_p)

    def check(A):
        A = 0.5 * (A + A.T)
        out = np.zeros((n, n))
        assert host_shim.hs_make_pd(n, P(np.ascontiguousarray(A)), P(out)) == 0
        lam, V = np.linalg.eigh(A)
        ref = (V * np.maximum(lam, 0.0)) @ V.T
        scale = max(np.abs(lam).max(), 1e-300)
        assert np.abs(out - ref).max() <= 1e-12 * scale, (np.abs(out - ref).max() / scale, lam)
        assert np.abs(out - out.T).max() == 0.0

    def with_spectrum(lam):
        Q, _ = np.linalg.qr(rng.normal(size=(n, n)))
        return (Q * np.asarray(lam, float)) @ Q.T

    for _ in range(300):
        check(rng.normal(size=(n, n)))
    check(np.zeros((n, n)))
    check(np.eye(n)); check(-np.eye(n))
    check(np.diag(np.arange(n) - n / 2.0))                                  # already diagonal
    v = rng.normal(size=n); check(np.outer(v, v)); check(-np.outer(v, v))    # rank one
    for _ in range(60):
        lam = rng.normal(size=n)
        lam[1] = lam[0]; lam[3] = lam[2] * (1 + 1e-13)                       # exact and near-exact pairs
        check(with_spectrum(lam))
        lam = np.concatenate([[-1.0] * (n // 2), [2.0] * (n - n // 2)])      # two clusters of multiplicity n/2
        check(with_spectrum(lam))
        lam = np.concatenate([[1e8], rng.normal(size=n - 3) * 1e-8, [0.0, -1e-8]])  # 1e16 range, exact zero
        check(with_spectrum(lam))
        check(with_spectrum(np.abs(rng.normal(size=n))))                     # definite: returned unchanged up to rounding
        check(with_spectrum(rng.normal(size=n)) * 1e-150); check(with_spectrum(rng.normal(size=n)) * 1e120)
    B = np.zeros((n, n)); B[:3, :3] = with_spectrum(rng.normal(size=n))[:3, :3]; B[3:, 3:] = with_spectrum(rng.normal(size=n))[3:, 3:]
    check(B)                                                                 # block diagonal: interior split from the start


def test_surface_primitive_ordering_contract():
    """boundaryEdge = lexicographic keys of a std::map with first-triangle orientation (Utils/MESHIO.h:768-834)."""
    from idp_b200 import meshgen
    V, F = meshgen.icosphere(5)
    assert len(V) == 10 * 25 + 2 and len(F) == 20 * 25
    m = meshgen.SurfaceMesh(V, F)
    assert len(m.bedge) == 30 * 25 and len(V) - len(m.bedge) + len(F) == 2  # Euler characteristic of a sphere
    assert np.array_equal(m.bnode, np.arange(len(V)))
    e = m.bedge.astype(np.int64)
    key = e[:, 0] * len(V) + e[:, 1]
    assert (np.diff(key) > 0).all()
    # restatement with a Python dict in the reference's visiting order
    ref = {}
    for a, b, c in F:
        for u, v in ((a, b), (b, c), (c, a)):
            if (v, u) not in ref:
                ref[(u, v)] = 1
    assert sorted(ref) == [tuple(x) for x in m.bedge]
    s, d = meshgen.sheet_stack(n_sheets=2, nx=4, ny=3)
    assert s.nF == 2 * 2 * 4 * 3 and s.nV == 2 * 5 * 4 and d.shape == s.X.shape
