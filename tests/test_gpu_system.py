"""GPU parity tests for the caller-side steps kept on the device, through the C ABI:
flow-term + lumped-mass blocks assembled with the barrier Hessian, Project_DBC, the PCG solve that stands in for
Solve_Direct, and the device-side surface-primitive extraction.

Bars: index / ordering outputs bit-exact; matrix values within 1e-10 relative (BASELINE.json's E/g/H tolerance); the
PCG solution within 1e-8 relative of a direct sparse solve (it is an iterative method: rel_tol = 1e-12 requested)."""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from conftest import make_cases

pytestmark = pytest.mark.gpu
KAPPA = 1e5
RTOL = 1e-10


def csr(ptr, col, val, n):
    A = sp.csr_matrix((val, col, ptr), shape=(n, n))
    A.sum_duplicates()
    return A


def rel_mat(A, B):
    d = (A - B)
    nb = spla.norm(B)
    return spla.norm(d) / nb if nb > 0 else spla.norm(d)


def tri_volumes(X, F, thickness=1e-3):
    a = 0.5 * np.linalg.norm(np.cross(X[F[:, 1]] - X[F[:, 0]], X[F[:, 2]] - X[F[:, 0]]), axis=1)
    return a * thickness


def lumped_mass(nV, X, F, thickness=1e-3, rho=1000.0):
    """Shell/DISCRETE_SHELL.h:279-318: massPortion = area * thickness * rho / 3 to each element vertex."""
    a = 0.5 * np.linalg.norm(np.cross(X[F[:, 1]] - X[F[:, 0]], X[F[:, 2]] - X[F[:, 0]]), axis=1)
    m = np.zeros(nV)
    np.add.at(m, F.reshape(-1), np.repeat(a * thickness * rho / 3.0, 3))
    return m


@pytest.fixture(scope="module")
def cases():
    return make_cases()


def test_system_matrix_flow_mass_dbc(lib_built, orc, cases):
    """[flow triplets][barrier triplets] -> Construct_From_Triplet -> += M -> Project_DBC (INC_POTENTIAL.h:321-394), oracle
    restatement vs the device assembly, with and without contact rows, with a Dirichlet mask."""
    from idp_b200 import ContactContext
    for name, m, _d, dhats in cases:
        rng = np.random.default_rng(3)
        dbc = (rng.uniform(size=m.nV) < 0.07).astype(np.uint8)
        F = np.ascontiguousarray(m.btri[:, :3], np.int32)
        vol = tri_volumes(m.X, F)
        mass = lumped_mass(m.nV, m.X, F)
        h = 0.01
        c = ContactContext(0)
        try:
            c.set_mesh(m.nV, m.bnode, m.bedge, m.btri, dbc)
            c.set_rest_positions(m.X0)
            c.set_positions(m.X)
            om = orc.mesh(m.X, m.X0, m.bnode, m.bedge, m.btri, dbc)
            for dh in (0.0, dhats[-1]):
                if dh > 0:
                    c.constraint_set(dh * dh)
                    rows, info = c.get_constraints()
                else:
                    rows = np.zeros((0, 4), np.int32)
                    c.set_constraints(rows)
                c.set_flow_term(F, vol, h)
                c.set_mass(mass)
                dh2 = max(dh, 1e-3) ** 2
                ptr, col, val = c.barrier_hessian(dh2, KAPPA, project_spd=True)
                optr, ocol, oval = orc.system_matrix(om, rows, np.ones(len(rows)), dh2, KAPPA, 0.0, True, F, vol, h, mass, False)
                A, B = csr(ptr, col, val, 3 * m.nV), csr(optr, ocol, oval, 3 * m.nV)
                assert rel_mat(A, B) <= RTOL, (name, dh, "system matrix", rel_mat(A, B))
                assert abs(A - A.T).max() <= 1e-12 * abs(A).max()
                # columns ascending inside every row (Eigen's setFromTriplets order)
                assert all(np.all(np.diff(col[ptr[r]:ptr[r + 1]]) > 0) for r in range(0, 3 * m.nV, 7))
                c.project_dbc()
                ptr2, col2, val2 = c.get_hessian_csr()
                assert np.array_equal(ptr, ptr2) and np.array_equal(col, col2)  # Project_DBC never changes the pattern
                optr, ocol, oval = orc.system_matrix(om, rows, np.ones(len(rows)), dh2, KAPPA, 0.0, True, F, vol, h, mass, True)
                A, B = csr(ptr2, col2, val2, 3 * m.nV), csr(optr, ocol, oval, 3 * m.nV)
                assert rel_mat(A, B) <= RTOL, (name, dh, "after Project_DBC")
                fixed = np.repeat(dbc.astype(bool), 3)
                D = A.toarray() if m.nV < 1500 else None
                if D is not None:
                    assert np.array_equal(D[fixed][:, fixed], np.eye(int(fixed.sum())))
                    assert not D[fixed][:, ~fixed].any() and not D[~fixed][:, fixed].any()
                # removing the terms again gives the bare barrier Hessian
                c.set_flow_term(None, None, 0.0)
                c.set_mass(None)
                if len(rows):
                    ptr3, col3, val3 = c.barrier_hessian(dh2, KAPPA, project_spd=True)
                    o3 = orc.barrier_hessian(om, rows, np.ones(len(rows)), dh2, KAPPA, project_spd=True)["csr"]
                    assert np.array_equal(ptr3, o3[0]) and np.array_equal(col3, o3[1])
        finally:
            c.close()


def test_pcg_solves_the_projected_system(lib_built, orc, cases):
    """idp_solve_pcg against a direct sparse solve of the same matrix (downloaded from the device)."""
    from idp_b200 import ContactContext
    name, m, _d, dhats = cases[1]
    rng = np.random.default_rng(5)
    dbc = (rng.uniform(size=m.nV) < 0.05).astype(np.uint8)
    F = np.ascontiguousarray(m.btri[:, :3], np.int32)
    c = ContactContext(0)
    try:
        c.set_mesh(m.nV, m.bnode, m.bedge, m.btri, dbc)
        c.set_rest_positions(m.X0)
        c.set_positions(m.X)
        c.constraint_set(dhats[-1] ** 2)
        c.set_flow_term(F, tri_volumes(m.X, F), 0.01)
        c.set_mass(lumped_mass(m.nV, m.X, F))
        c.barrier_hessian(dhats[-1] ** 2, KAPPA, project_spd=True, fetch=False)
        c.project_dbc()
        ptr, col, val = c.get_hessian_csr()
        A = csr(ptr, col, val, 3 * m.nV).tocsc()
        rhs = rng.normal(size=3 * m.nV)
        ref = spla.spsolve(A, rhs)
        sol, iters, res = c.solve_pcg(rhs, rel_tol=1e-12, max_iter=20000)
        assert res <= 1e-12 and iters > 0, (iters, res)
        assert np.linalg.norm(sol - ref) <= 1e-8 * np.linalg.norm(ref), (iters, res, np.linalg.norm(sol - ref) / np.linalg.norm(ref))
        # the TRUE residual drifts from the recurrence residual PCG monitors (ill-conditioned barrier stiffness): 1e-8 is the bar
        assert np.linalg.norm(A @ sol - rhs) <= 1e-8 * np.linalg.norm(rhs)
        # a second right-hand side on the same matrix, looser tolerance, fewer iterations
        sol2, it2, res2 = c.solve_pcg(rhs, rel_tol=1e-6, max_iter=20000)
        assert it2 <= iters and res2 <= 1e-6
        # error behaviour: no matrix -> IDP_ERR_INVALID
        c2 = ContactContext(0)
        c2.set_mesh(m.nV, m.bnode, m.bedge, m.btri, None)
        with pytest.raises(Exception):
            c2.solve_pcg(rhs)
        c2.close()
    finally:
        c.close()


def test_surface_primitives_on_the_device(lib_built, orc):
    """idp_set_mesh_from_triangles against the std::map restatement of MESHIO.h:768-834: closed and open meshes, a mesh with an
    unreferenced vertex, a zero-area triangle, and a soup with inconsistent orientation and a non-manifold edge."""
    from idp_b200 import ContactContext, meshgen
    meshes = []
    X, F = meshgen.icosphere(6)
    meshes.append(("icosphere", X, F))
    ms, _ = meshgen.sheet_stack(n_sheets=2, nx=17, ny=9, h=0.05, A=0.01)
    meshes.append(("sheets", ms.X, np.ascontiguousarray(ms.btri[:, :3])))
    rng = np.random.default_rng(2)
    Xs = rng.uniform(-1, 1, (40, 3))
    Fs = np.array([rng.choice(39, 3, replace=False) for _ in range(120)], np.int32)  # vertex 39 unreferenced
    Fs[5] = Fs[4][[1, 0, 2]]            # same undirected edges as triangle 4, opposite orientation
    Fs[7, :2] = Fs[4, :2]               # third triangle on the edge (Fs[4,0], Fs[4,1]): non-manifold
    Xs[Fs[9, 2]] = Xs[Fs[9, 1]]         # zero-area triangle
    meshes.append(("soup", Xs, Fs))
    c = ContactContext(0)
    try:
        for name, X, F in meshes:
            c.set_mesh_from_triangles(len(X), F, X)
            got = c.get_surface_primitives(areas=True)
            want = orc.surface(len(X), F, X)
            for k in ("bnode", "bedge", "btri"):
                assert np.array_equal(got[k], want[k]), (name, k)
            for k in ("BNArea", "BEArea", "BTArea"):
                assert np.allclose(got[k], want[k], rtol=1e-13, atol=1e-300), (name, k)
            # the context is usable as if idp_set_mesh had been called with these arrays
            c.set_rest_positions(X)
            c.set_positions(X)
            n = c.constraint_set(1e-4)
            c2 = ContactContext(0)
            c2.set_mesh(len(X), want["bnode"], want["bedge"], want["btri"], None)
            c2.set_rest_positions(X)
            c2.set_positions(X)
            assert n == c2.constraint_set(1e-4)
            assert np.array_equal(c.get_constraints()[0], c2.get_constraints()[0])
            c2.close()
    finally:
        c.close()
