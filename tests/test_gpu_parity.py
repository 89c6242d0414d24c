"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on identical seeded inputs.

Bars (BASELINE.json): sorted candidate and constraint sets bit-exact; barrier E/g/H within 1e-10 relative;
CCD step never above the oracle's and within 1e-6 relative of it.
"""
import numpy as np
import pytest

from conftest import lexsorted, make_cases

pytestmark = pytest.mark.gpu
KAPPA = 1e5
RTOL = 1e-10  # tolerance stated by BASELINE.json for E / g / H


def rel(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    nb = np.linalg.norm(b)
    return np.linalg.norm(a - b) / nb if nb > 0 else np.linalg.norm(a - b)


@pytest.fixture(scope="module")
def cases():
    return make_cases()


def omesh(orc, m):
    return orc.mesh(m.X, m.X0, m.bnode, m.bedge, m.btri, m.dbc)


def test_constraint_and_candidate_sets_bit_exact(gpu_ctx, orc, cases):
    for name, m, _d, dhats in cases:
        gpu_ctx.set_surface_mesh(m)
        om = omesh(orc, m)
        for dh in dhats:
            n = gpu_ctx.constraint_set(dh * dh)
            rows, info = gpu_ctx.get_constraints()
            cpt, cee = gpu_ctx.get_candidates(0), gpu_ctx.get_candidates(1)
            orows, oinfo, ocpt, ocee = orc.constraint_set(om, dh * dh, want_cand=True)
            assert n == len(orows), (name, dh, n, len(orows))
            assert np.array_equal(cpt, ocpt), (name, dh, "PT candidates")
            assert np.array_equal(cee, ocee), (name, dh, "EE candidates")
            assert np.array_equal(lexsorted(rows), lexsorted(orows)), (name, dh, "rows")
            assert np.array_equal(info, np.tile([1.0, oinfo[0, 1] if len(oinfo) else dh * dh], (n, 1)))
            # merged PP/PE group is in key order at the tail, exactly as the reference emits it (IPC.h:651-654)
            dup = (rows[:, 0] < 0) & (rows[:, 3] < 0)
            odup = (orows[:, 0] < 0) & (orows[:, 3] < 0)
            assert np.array_equal(rows[dup], orows[odup]), (name, dh, "merged group order")


def test_constraint_set_with_thickness_and_dbc(gpu_ctx, orc, cases):
    name, m, _d, dhats = cases[0]
    dbc = np.zeros(m.nV, np.uint8)
    dbc[: m.nV // 2] = 1  # the whole inner sphere is Dirichlet: its self-pairs must vanish
    gpu_ctx.set_mesh(m.nV, m.bnode, m.bedge, m.btri, dbc)
    gpu_ctx.set_rest_positions(m.X0)
    gpu_ctx.set_positions(m.X)
    om = orc.mesh(m.X, m.X0, m.bnode, m.bedge, m.btri, dbc)
    for thickness in (0.0, 5e-3):
        n = gpu_ctx.constraint_set(dhats[1] ** 2, thickness)
        rows, info = gpu_ctx.get_constraints()
        orows, oinfo, ocpt, ocee = orc.constraint_set(om, dhats[1] ** 2, thickness, want_cand=True)
        assert n == len(orows) and n > 0
        assert np.array_equal(lexsorted(rows), lexsorted(orows))
        assert np.array_equal(gpu_ctx.get_candidates(0), ocpt) and np.array_equal(gpu_ctx.get_candidates(1), ocee)
        assert info[0, 1] == oinfo[0, 1]


def test_barrier_energy_gradient_hessian(gpu_ctx, orc, cases):
    for name, m, _d, dhats in cases:
        gpu_ctx.set_surface_mesh(m)
        om = omesh(orc, m)
        dh = dhats[-1]
        n = gpu_ctx.constraint_set(dh * dh)
        assert n > 0
        rows, info = gpu_ctx.get_constraints()
        for project in (False, True):
            E = gpu_ctx.barrier_energy(dh * dh, KAPPA, E0=1.5)
            g = gpu_ctx.barrier_gradient(dh * dh, KAPPA)
            ptr, col, val = gpu_ctx.barrier_hessian(dh * dh, KAPPA, project_spd=project)
            st, oE = orc.barrier(om, rows, info[:, 0], dh * dh, KAPPA)
            st2, og = orc.barrier_gradient(om, rows, info[:, 0], dh * dh, KAPPA)
            oh = orc.barrier_hessian(om, rows, info[:, 0], dh * dh, KAPPA, project_spd=project)
            assert st == 0 and st2 == 0 and oh["status"] == 0
            assert abs((E - 1.5) - oE) <= RTOL * abs(oE), (name, E, oE)
            assert rel(g, og) <= RTOL, (name, rel(g, og))
            optr, ocol, oval = oh["csr"]
            assert np.array_equal(ptr, optr) and np.array_equal(col, ocol), (name, "CSR pattern")
            assert rel(val, oval) <= RTOL, (name, project, rel(val, oval))
            assert np.abs(val - oval).max() <= RTOL * np.abs(oval).max()


def test_cuda_path_matches_reference_golden(gpu_ctx, cases):
    """Directly against what the REFERENCE's own loops returned (tests/golden/ref_loops.npz, produced from FEM/IPC.h and
    Grid/SPATIAL_HASH.h compiled from /root/reference; see tests/test_ref_loops.py), without the oracle in between."""
    import hashlib
    import os
    import scipy.sparse as sp
    from conftest import ROOT
    G = np.load(os.path.join(ROOT, "tests", "golden", "ref_loops.npz"))
    sha = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()
    for name, m, d, dhats in cases:
        gpu_ctx.set_surface_mesh(m)
        for k, dh in enumerate(dhats):
            n = gpu_ctx.constraint_set(dh * dh)
            rows, info = gpu_ctx.get_constraints()
            assert n == int(G["%s/cs%d/n" % (name, k)]), (name, dh)
            assert sha(lexsorted(rows).astype(np.int32)) == str(G["%s/cs%d/sorted_sha" % (name, k)]), (name, dh)
            dup = (rows[:, 0] < 0) & (rows[:, 3] < 0)
            assert sha(rows[dup].astype(np.int32)) == str(G["%s/cs%d/merged_sha" % (name, k)]), (name, dh)
            if n:
                assert np.array_equal(info[0], G["%s/cs%d/info" % (name, k)])
        dh = dhats[-1]
        grows = G["%s/rows" % name]
        gpu_ctx.set_constraints(grows)
        E = gpu_ctx.barrier_energy(dh * dh, KAPPA)
        g = gpu_ctx.barrier_gradient(dh * dh, KAPPA)
        assert abs(E - float(G["%s/E" % name])) <= RTOL * abs(float(G["%s/E" % name]))
        assert rel(g, G["%s/g" % name]) <= RTOL
        N = 3 * m.nV
        probe = np.random.default_rng(20260118).normal(size=N)
        for spd in (0, 1):
            ptr, col, val = gpu_ctx.barrier_hessian(dh * dh, KAPPA, project_spd=bool(spd))
            H = sp.csr_matrix((val, col, ptr), shape=(N, N))
            assert rel(H @ probe, G["%s/H%d_probe" % (name, spd)]) <= RTOL, (name, spd)
            assert abs(np.sqrt((val ** 2).sum()) - float(G["%s/H%d_fro" % (name, spd)])) <= RTOL * float(G["%s/H%d_fro" % (name, spd)])
            assert H.nnz == int(G["%s/H%d_nnz" % (name, spd)])
        d2, mn = gpu_ctx.min_dist2(thickness=1e-4)
        assert sha(d2) == str(G["%s/dist2_sha" % name]) and mn == float(G["%s/min_dist2" % name])
        for (scale, xi, a0), ref_step in zip(((1.0, 0.0, 1.0), (0.3, 0.0, 1.0), (4.0, 0.0, 1.0), (1.0, 1e-4, 0.7)), G["%s/ccd" % name]):
            a = gpu_ctx.ccd_step(d * scale, a0, xi)
            assert a <= ref_step and abs(a - ref_step) <= 1e-6 * ref_step, (name, scale, xi, a, ref_step)


@pytest.mark.parametrize("name", ["bunny3K", "hand"])
def test_cuda_path_matches_reference_on_paper_meshes(gpu_ctx, name):
    """Real geometry of the paper examples against what the reference's own loops returned (tests/golden/
    ref_paper_meshes.npz; configuration and generator: tests/golden/make_golden_paper.py)."""
    import hashlib
    import os
    import scipy.sparse as sp
    from conftest import ROOT
    from test_ref_loops import paper_case
    P = np.load(os.path.join(ROOT, "tests", "golden", "ref_paper_meshes.npz"))
    sha = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()
    m, d, e = paper_case(P, name)
    gpu_ctx.set_surface_mesh(m)
    for k, f in enumerate((0.6, 1.2)):
        dh = f * e
        n = gpu_ctx.constraint_set(dh * dh)
        rows, info = gpu_ctx.get_constraints()
        assert n == int(P["%s/cs%d/n" % (name, k)])
        assert sha(lexsorted(rows).astype(np.int32)) == str(P["%s/cs%d/sorted_sha" % (name, k)])
        dup = (rows[:, 0] < 0) & (rows[:, 3] < 0)
        assert sha(rows[dup].astype(np.int32)) == str(P["%s/cs%d/merged_sha" % (name, k)])
    E = gpu_ctx.barrier_energy(dh * dh, KAPPA)
    g = gpu_ctx.barrier_gradient(dh * dh, KAPPA)
    ptr, col, val = gpu_ctx.barrier_hessian(dh * dh, KAPPA, project_spd=True)
    assert abs(E - float(P[name + "/E"])) <= RTOL * abs(float(P[name + "/E"])) and rel(g, P[name + "/g"]) <= RTOL
    N = 3 * m.nV
    probe = np.random.default_rng(20260118).normal(size=N)
    assert rel(sp.csr_matrix((val, col, ptr), shape=(N, N)) @ probe, P[name + "/H_probe"]) <= RTOL
    _, mn = gpu_ctx.min_dist2()
    assert mn == float(P[name + "/min_dist2"])
    for (scale, a0), ref_step in zip(((1.0, 1.0), (0.25, 0.5)), P[name + "/ccd"]):
        a = gpu_ctx.ccd_step(d * scale, a0, 0.0)
        assert a <= ref_step and abs(a - ref_step) <= 1e-6 * ref_step, (name, a, ref_step)


def test_barrier_all_matches_separate_calls(gpu_ctx, cases):
    name, m, _d, dhats = cases[1]
    gpu_ctx.set_surface_mesh(m)
    dh = dhats[-1]
    gpu_ctx.constraint_set(dh * dh)
    E1 = gpu_ctx.barrier_energy(dh * dh, KAPPA)
    ptr1, col1, val1 = gpu_ctx.barrier_hessian(dh * dh, KAPPA)
    E2, nnz = gpu_ctx.barrier_all(dh * dh, KAPPA)
    ptr2, col2, val2 = gpu_ctx.get_hessian_csr()
    assert abs(E1 - E2) <= 1e-12 * abs(E1) and nnz == len(col1)
    assert np.array_equal(ptr1, ptr2) and np.array_equal(col1, col2) and np.allclose(val1, val2, rtol=1e-11, atol=1e-11 * np.abs(val1).max())


def test_min_dist2(gpu_ctx, orc, cases):
    for name, m, _d, dhats in cases:
        gpu_ctx.set_surface_mesh(m)
        om = omesh(orc, m)
        gpu_ctx.constraint_set(dhats[-1] ** 2)
        rows, _ = gpu_ctx.get_constraints()
        d, mn = gpu_ctx.min_dist2(thickness=1e-4)
        od, omn = orc.min_dist2(om, rows, thickness=1e-4)
        assert np.array_equal(d, od) and mn == omn, name


def test_ccd_step_conservative_and_close(gpu_ctx, orc, cases):
    for name, m, d, _dh in cases:
        gpu_ctx.set_surface_mesh(m)
        om = omesh(orc, m)
        for scale, thickness, a0 in ((1.0, 0.0, 1.0), (0.3, 0.0, 1.0), (4.0, 0.0, 1.0), (1.0, 1e-4, 0.7)):
            dd = d * scale
            a = gpu_ctx.ccd_step(dd, a0, thickness)
            o = orc.ccd(om, dd, a0, thickness, want_cand=True)
            assert o["status"] == 0
            assert a <= o["step"], (name, scale, a, o["step"])
            assert abs(a - o["step"]) <= 1e-6 * o["step"], (name, scale, a, o["step"])
            clamped = o["step_after_clamp"] != a0
            if not clamped:
                # same lattice on both sides -> candidate sets reaching the ACCD calls are bit-exact
                assert np.array_equal(gpu_ctx.get_candidates(2), o["cand_pt"]), (name, scale, "CCD PT candidates")
                assert np.array_equal(gpu_ctx.get_candidates(3), o["cand_ee"]), (name, scale, "CCD EE candidates")
                assert a == o["step"] or o["step"] == a0 or a == o["step"]


def test_ccd_result_is_intersection_free(gpu_ctx, orc, cases):
    """Size-independent safety property: every constraint distance at x + alpha*dir stays positive."""
    name, m, d, _dh = cases[0]
    gpu_ctx.set_surface_mesh(m)
    a = gpu_ctx.ccd_step(d * 2.0, 1.0, 0.0)
    assert 0 < a < 1
    X1 = m.X + a * d * 2.0
    gpu_ctx.set_positions(X1)
    n = gpu_ctx.constraint_set(0.05 ** 2)
    assert n > 0
    dist, mn = gpu_ctx.min_dist2()
    assert mn > 0
    gpu_ctx.set_positions(m.X)


def test_nonpositive_distance_is_reported(gpu_ctx):
    from idp_b200 import IdpError
    X = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0.2, 0.2, 0.0]], np.float64)  # point in the triangle's plane
    gpu_ctx.set_mesh(4, np.arange(4, dtype=np.int32), np.array([[0, 1], [1, 2], [2, 0]], np.int32), np.array([[0, 1, 2]], np.int32))
    gpu_ctx.set_rest_positions(X)
    gpu_ctx.set_positions(X)
    gpu_ctx.set_constraints(np.array([[-4, 0, 1, 2]], np.int32))
    with pytest.raises(IdpError) as e:
        gpu_ctx.barrier_energy(1e-2, KAPPA)
    assert e.value.code == 3


def test_unsupported_inputs_rejected(gpu_ctx):
    from idp_b200 import IdpError
    with pytest.raises(IdpError) as e:
        gpu_ctx.declare_unsupported(n_rod=1)
    assert e.value.code == 5
    gpu_ctx.declare_unsupported()


def test_empty_constraint_set(gpu_ctx, cases):
    name, m, d, _ = cases[0]
    gpu_ctx.set_surface_mesh(m)
    assert gpu_ctx.constraint_set(1e-12) == 0
    assert gpu_ctx.barrier_energy(1e-12, KAPPA) == 0.0
    ptr, col, val = gpu_ctx.barrier_hessian(1e-12, KAPPA)
    assert len(col) == 0 and not ptr.any()
    assert not gpu_ctx.barrier_gradient(1e-12, KAPPA).any()


def test_b2_dropin_operators(lib_built, orc, cases):
    """The six operators of idp_b200/host/IPC_B200.h (reference signatures, mock JGSL containers) against the oracle."""
    import ctypes as C
    import os
    import subprocess
    from conftest import ROOT
    src = os.path.join(ROOT, "tests", "host_shim", "b2_driver.cpp")
    out = os.path.join(ROOT, "tests", "host_shim", "libb2_driver.so")
    subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-o", out, src, "-L", os.path.join(ROOT, "idp_b200"),
                           "-lidp_contact", "-Wl,-rpath," + os.path.join(ROOT, "idp_b200")])
    drv = C.CDLL(out)
    name, m, d, dhats = cases[1]
    dh, thickness = dhats[-1], 0.0
    P = lambda a: a.ctypes.data_as(C.c_void_p)
    cap = 400000
    rows = np.zeros((cap, 4), np.int32); g = np.zeros((m.nV, 3)); dist2 = np.zeros(cap)
    tcap = 40000000
    tr = np.zeros(tcap, np.int32); tc = np.zeros(tcap, np.int32); tv = np.zeros(tcap)
    n = C.c_int(0); E = C.c_double(0); nt = C.c_long(0); alpha = C.c_double(0); mn = C.c_double(0)
    drv.b2_run.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                           C.c_double, C.c_double, C.c_double, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                           C.c_void_p, C.c_void_p, C.c_void_p, C.c_long, C.c_void_p, C.c_void_p, C.c_void_p]
    X = np.ascontiguousarray(m.X); X0 = np.ascontiguousarray(m.X0); dd = np.ascontiguousarray(d)
    st = drv.b2_run(m.nV, P(X), P(X0), len(m.bnode), P(m.bnode), len(m.bedge), P(m.bedge), len(m.btri), P(m.btri), P(dd),
                    dh * dh, KAPPA, thickness, C.byref(n), P(rows), cap, C.byref(E), P(g), C.byref(nt), P(tr), P(tc), P(tv), tcap,
                    C.byref(alpha), P(dist2), C.byref(mn))
    assert st == 0 and 0 < n.value < cap and nt.value < tcap
    rows = rows[: n.value]
    om = omesh(orc, m)
    orows, oinfo, _, _ = orc.constraint_set(om, dh * dh)
    assert np.array_equal(lexsorted(rows), lexsorted(orows))
    w = np.ones(len(rows))
    _, oE = orc.barrier(om, rows, w, dh * dh, KAPPA)
    _, og = orc.barrier_gradient(om, rows, w, dh * dh, KAPPA)
    assert abs((E.value - 0.25) - oE) <= RTOL * abs(oE) and rel(g, og) <= RTOL
    import scipy.sparse as sp
    A = sp.coo_matrix((tv[1:nt.value], (tr[1:nt.value], tc[1:nt.value])), shape=(3 * m.nV, 3 * m.nV)).tocsr()
    assert (tr[0], tc[0], tv[0]) == (0, 0, 1.0)  # appended, not overwritten
    optr, ocol, oval = orc.barrier_hessian(om, rows, w, dh * dh, KAPPA)["csr"]
    B = sp.csr_matrix((oval, ocol, optr), shape=A.shape)
    assert abs(A - B).max() <= RTOL * abs(B).max()
    oc = orc.ccd(om, dd, 1.0, thickness)
    assert alpha.value <= oc["step"] and abs(alpha.value - oc["step"]) <= 1e-6 * oc["step"]
    od, omn = orc.min_dist2(om, rows, thickness)
    assert np.array_equal(dist2[: n.value], od) and mn.value == omn


def test_b2_side_by_side_with_reference_types(lib_built, cases):
    """tests/host_shim/b2_side_by_side.cpp: the reference's own FEM/IPC.h operators and JGSL::B200::Compute_* called with
    the SAME argument objects (reference VECTOR / storage stand-ins / Eigen::Triplet) in one translation unit."""
    import ctypes as C
    import os
    from conftest import ROOT
    so = os.path.join(ROOT, "tests", "host_shim", "libb2_side_by_side.so")
    if not os.path.exists(so):
        pytest.skip("libb2_side_by_side.so is built where /root/reference exists (make -C tests/host_shim)")
    drv = C.CDLL(so)
    drv.b2_side_by_side.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                    C.c_double, C.c_double, C.c_double, C.c_void_p]
    P = lambda a: a.ctypes.data_as(C.c_void_p)
    for name, m, d, dhats in cases[:2]:
        dh = dhats[-1]
        X = np.ascontiguousarray(m.X); X0 = np.ascontiguousarray(m.X0); dd = np.ascontiguousarray(d)
        rep = np.zeros(16)
        assert drv.b2_side_by_side(m.nV, P(X), P(X0), len(m.bnode), P(m.bnode), len(m.bedge), P(m.bedge), len(m.btri), P(m.btri), P(dd),
                                   dh * dh, KAPPA, 0.0, P(rep)) == 0
        n_ref, n_new, same, e_ref, e_new, gmax, gdiff, hmax, hdiff, pattern, nnz, a_ref, a_new, m_ref, m_new, deq = rep
        assert n_ref == n_new > 0 and same == 1.0, (name, n_ref, n_new)
        assert abs(e_new - e_ref) <= RTOL * abs(e_ref - 0.25)          # E is accumulated onto the caller's value (0.25)
        assert gdiff <= RTOL * gmax and hdiff <= RTOL * hmax and pattern == 1.0 and nnz > 0
        assert a_new <= a_ref and abs(a_new - a_ref) <= 1e-6 * a_ref
        assert m_new == m_ref and deq == 1.0


def test_async_transfers_match_the_synchronous_getters(gpu_ctx, cases):
    """idp_get_constraints_begin / idp_get_hessian_csr_begin + idp_transfers_end (copy stream, compact block-column CSR
    expanded on the host) against idp_get_constraints / idp_get_hessian_csr, with other operators running in between."""
    c = gpu_ctx
    for name, m, d, dhats in cases[:2]:
        c.set_surface_mesh(m)
        dh = dhats[-1]
        n = c.constraint_set(dh * dh)
        rows, info = c.get_constraints()
        rows2 = np.full_like(rows, -7); info2 = np.full_like(info, -7.0)
        c._ck(c.L.idp_get_constraints_begin(c.h, rows2.ctypes.data, info2.ctypes.data))
        E, nnz = c.barrier_all(dh * dh, KAPPA)            # runs while the rows travel
        ptr = np.full(3 * m.nV + 1, -7, np.int32); col = np.full(nnz, -7, np.int32); val = np.full(nnz, -7.0)
        c._ck(c.L.idp_get_hessian_csr_begin(c.h, ptr.ctypes.data, col.ctypes.data, val.ctypes.data))
        a = c.ccd_step(d, 1.0)                             # runs while the CSR travels
        c.min_dist2(want_all=False)
        c._ck(c.L.idp_transfers_end(c.h))
        assert np.array_equal(rows, rows2) and np.array_equal(info, info2), name
        sptr, scol, sval = c.get_hessian_csr()
        assert np.array_equal(ptr, sptr) and np.array_equal(col, scol) and np.array_equal(val, sval), name
        # a second begin without end is refused; end with nothing pending is a no-op
        c._ck(c.L.idp_transfers_end(c.h))
        c._ck(c.L.idp_get_hessian_csr_begin(c.h, ptr.ctypes.data, col.ctypes.data, val.ctypes.data))
        assert c.L.idp_get_hessian_csr_begin(c.h, ptr.ctypes.data, col.ctypes.data, val.ctypes.data) != 0
        c._ck(c.L.idp_transfers_end(c.h))


def test_b2_flow_system_matrix_with_reference_types(lib_built, orc, cases):
    """tests/host_shim/b2_side_by_side.cpp::b2_flow_system: the `flow` branch of Compute_IncPotential_Hessian
    (Shell/INC_POTENTIAL.h:321-394) -- the reference's own Compute_Barrier_Hessian + CSR_MATRIX (Construct_From_Triplet,
    += M, Project_DBC) against B200::Compute_IncPotential_Hessian_Flow (device assembly, Construct_From_CSR hand-over)."""
    import ctypes as C
    import os
    from conftest import ROOT
    so = os.path.join(ROOT, "tests", "host_shim", "libb2_side_by_side.so")
    if not os.path.exists(so):
        pytest.skip("libb2_side_by_side.so is built where /root/reference exists (make -C tests/host_shim)")
    drv = C.CDLL(so)
    if not hasattr(drv, "b2_flow_system"):
        pytest.skip("stale libb2_side_by_side.so")
    drv.b2_flow_system.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p,
                                   C.c_double, C.c_double, C.c_double, C.c_void_p]
    P = lambda a: a.ctypes.data_as(C.c_void_p)
    for name, m, d, dhats in cases[:2]:
        dh = dhats[-1]
        rng = np.random.default_rng(8)
        rows, _info, _, _ = orc.constraint_set(omesh(orc, m), dh * dh)
        rows = np.ascontiguousarray(rows, np.int32)
        F = np.ascontiguousarray(m.btri[:, :3], np.int32)
        vol = rng.uniform(0.5, 2.0, len(F)) * 1e-4
        mass = rng.uniform(0.1, 1.0, m.nV)
        dbc = (rng.uniform(size=m.nV) < 0.06).astype(np.uint8)
        X = np.ascontiguousarray(m.X); X0 = np.ascontiguousarray(m.X0)
        rep = np.zeros(8)
        assert drv.b2_flow_system(m.nV, P(X), P(X0), len(F), P(F), P(vol), P(mass), P(dbc), len(rows), P(rows), 0.01, dh * dh, KAPPA, P(rep)) == 0
        nnz_ref, nnz_new, amax, dmax, missing = rep[:5]
        assert nnz_ref > 0 and nnz_new >= nnz_ref and missing == 0, (name, rep[:5])
        assert dmax <= RTOL * amax, (name, dmax, amax)


def test_ipc_energy_plugin_routed_to_b200(lib_built, orc, cases):
    """tests/host_shim/ipc_energy_plugin.cpp: the reference's own plugin class IPC_ENERGY<T,3,false> (FEM/Energy/IPC_ENERGY.h:11-58,
    compiled unmodified from /root/reference) with its three barrier calls qualified B200::, driven through the
    ABSTRACT_ENERGY virtual interface, against the reference's free functions on the same objects."""
    import ctypes as C
    import os
    from conftest import ROOT
    so = os.path.join(ROOT, "tests", "host_shim", "libipc_energy_plugin.so")
    if not os.path.exists(so):
        pytest.skip("libipc_energy_plugin.so is built where /root/reference exists (make -C tests/host_shim)")
    drv = C.CDLL(so)
    drv.ipc_energy_plugin.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_double, C.c_double, C.c_void_p]
    P = lambda a: a.ctypes.data_as(C.c_void_p)
    for name, m, d, dhats in cases[:2]:
        dh = dhats[-1]
        rows, _info, _, _ = orc.constraint_set(omesh(orc, m), dh * dh)
        rows = np.ascontiguousarray(rows, np.int32)
        X = np.ascontiguousarray(m.X); X0 = np.ascontiguousarray(m.X0)
        rep = np.zeros(8)
        assert drv.ipc_energy_plugin(m.nV, P(X), P(X0), len(rows), P(rows), dh * dh, KAPPA, P(rep)) == 0
        e_ref, e_new, gmax, gdiff, hmax, hdiff, n_ref, n_new = rep
        assert len(rows) > 0 and abs(e_new - e_ref) <= RTOL * abs(e_ref - 0.5), (name, e_ref, e_new)   # added onto the caller's 0.5
        assert gmax > 0 and gdiff <= RTOL * gmax, (name, gdiff, gmax)
        assert n_new == n_ref > 0 and hdiff <= RTOL * hmax, (name, n_ref, n_new, hdiff, hmax)


def test_row_merge_fallback_for_large_vertex_counts(lib_built, orc, cases, monkeypatch):
    """Meshes with more than 2^21 vertices cannot pack a PP/PE row into one 64-bit key; the 16-byte row merge sort they fall
    back to is forced here on a small mesh (IDP_FORCE_ROW_MERGE) and must give the identical constraint set."""
    from idp_b200 import ContactContext
    monkeypatch.setenv("IDP_FORCE_ROW_MERGE", "1")
    ctx = ContactContext(0)
    try:
        for name, m, d, dhats in cases[:2]:
            ctx.set_surface_mesh(m)
            om = omesh(orc, m)
            n = ctx.constraint_set(dhats[-1] ** 2)
            rows, _ = ctx.get_constraints()
            orows, _, _, _ = orc.constraint_set(om, dhats[-1] ** 2)
            assert n == len(orows) and np.array_equal(lexsorted(rows), lexsorted(orows)), name
            dup = (rows[:, 0] < 0) & (rows[:, 3] < 0)
            assert np.array_equal(rows[dup], orows[(orows[:, 0] < 0) & (orows[:, 3] < 0)]), name
    finally:
        ctx.close()


def test_hessian_assembly_bucket_sizes(lib_built, orc):
    """The CSR assembly reduces the 3x3 blocks per lower vertex: buckets of <= 128 / 256 / 512 entries are sorted in
    registers by one warp, larger ones by one CTA in global scratch. A 'star' of point-point and point-edge rows around a few
    hub vertices drives every path (hub 0: ~4000 entries, hubs 1-3: 100-500) and is compared with the oracle's CSR; the
    result must also be bit-identical from run to run (duplicates are summed in origin order, not in arrival order)."""
    from idp_b200 import ContactContext
    rng = np.random.default_rng(17)
    nV = 3000
    X = rng.uniform(-1, 1, (nV, 3))
    X[:4] = [[0, 0, 0], [5, 0, 0], [0, 5, 0], [0, 0, 5]]
    rows = []
    for hub, cnt in ((0, 2000), (1, 230), (2, 120), (3, 50)):
        partners = rng.choice(np.arange(4, nV), cnt, replace=False)
        X[partners] = X[hub] + rng.normal(0, 1, (cnt, 3)) * 0.02 + 0.05 * rng.choice([-1, 1], (cnt, 3))
        for k, b in enumerate(partners):
            if k % 3 == 0 and k + 1 < cnt:
                rows.append((-hub - 1, int(b), int(partners[k + 1]), -1 - (k % 2)))   # point-edge, multiplicity 1 or 2
            else:
                rows.append((-hub - 1, int(b), -1, -1 - (k % 3)))                      # point-point, multiplicity 1..3
    rows = np.array(rows, np.int32)
    dhat2 = 0.5 ** 2
    c = ContactContext(0)
    c.set_mesh(nV, np.arange(nV, dtype=np.int32), np.zeros((0, 2), np.int32), np.zeros((0, 3), np.int32))
    c.set_rest_positions(X)
    c.set_positions(X)
    c.set_constraints(rows)
    ptr, col, val = c.barrier_hessian(dhat2, KAPPA, project_spd=True)
    ptr2, col2, val2 = c.barrier_hessian(dhat2, KAPPA, project_spd=True)
    assert np.array_equal(ptr, ptr2) and np.array_equal(col, col2) and np.array_equal(val, val2)
    om = orc.mesh(X, X, np.arange(nV, dtype=np.int32), np.zeros((0, 2), np.int32), np.zeros((0, 3), np.int32))
    optr, ocol, oval = orc.barrier_hessian(om, rows, np.ones(len(rows)), dhat2, KAPPA, project_spd=True)["csr"]
    assert np.array_equal(ptr, optr) and np.array_equal(col, ocol)
    assert rel(val, oval) <= RTOL and np.abs(val - oval).max() <= RTOL * np.abs(oval).max()
    c.close()


def test_config2_geometry_matches_reference_loops(lib_built, orc):
    """BASELINE configs[1] geometry (wm2_15k, 12,811 vertices / 25,472 triangles, dHat = 1e-2; tests/golden/fix_char_seq_trace.npz): the
    contact-rich, intersection-free state the animation-fix example reaches after six frames of Rumba_Dancing_unfixed (46.7 K rows),
    search direction = towards the (self-intersecting) target frame. The hot-path operators on the paper's own mesh against the
    reference's loops (oracle/_ref) where that build exists, else against the oracle: constraint set bit-exact, E / g 1e-10,
    min-dist bit-exact, CCD conservative and within 1e-6."""
    import os
    from idp_b200 import ContactContext, meshgen
    from oracle import ref_binding
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fix_char_seq_trace.npz")
    if not os.path.exists(path):
        pytest.skip("fix_char_seq fixture absent")
    z = np.load(path)
    F = z["rest/F"]
    X, Xnext = z["V_end"], z["frame6/V"]
    m = meshgen.SurfaceMesh(X, F, X0=z["frame6/V"])
    dh2 = 1e-4
    c = ContactContext(0)
    try:
        c.set_surface_mesh(m)
        n = c.constraint_set(dh2)
        rows, info = c.get_constraints()
        assert n > 10000
        if ref_binding.ipc_available():
            ref = ref_binding.ReferenceIPC()
            rrows, rinfo = ref.constraint_set(m, dh2)
            rE, rg, _ = ref.barrier(m, rows, info[:, 0], dh2, KAPPA, want_h=False)
            rd, rmin = ref.min_dist2(m, rows)
            ralpha = ref.ccd(m, Xnext - X, 1.0)
        else:
            om = orc.mesh(m.X, m.X0, m.bnode, m.bedge, m.btri, m.dbc)
            rrows, rinfo, _, _ = orc.constraint_set(om, dh2)
            _, rE = orc.barrier(om, rows, info[:, 0], dh2, KAPPA)
            rg = orc.barrier_gradient(om, rows, info[:, 0], dh2, KAPPA)[1]
            rd, rmin = orc.min_dist2(om, rows)
            ralpha = orc.ccd(om, Xnext - X, 1.0)["step"]
        assert n == len(rrows) and np.array_equal(lexsorted(rows), lexsorted(rrows))
        E = c.barrier_energy(dh2, KAPPA)
        g = c.barrier_gradient(dh2, KAPPA)
        assert abs(E - rE) <= RTOL * abs(rE)
        assert np.abs(g - rg).max() <= RTOL * np.abs(rg).max()
        d, mn = c.min_dist2()
        assert np.array_equal(d, rd) and mn == rmin
        alpha = c.ccd_step(Xnext - X, 1.0)
        assert alpha <= ralpha and abs(alpha - ralpha) <= 1e-6 * ralpha, (alpha, ralpha)
    finally:
        c.close()
