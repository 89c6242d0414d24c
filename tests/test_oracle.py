"""CPU tests of the oracle (oracle/): self-consistency checks that pin the restatement, since the reference ships no
golden vectors for this path. Golden vectors produced by the reference's own generated derivative
code (through oracle/_ref) are checked in tests/test_golden.py."""
import numpy as np
import pytest

from conftest import lexsorted, make_cases

RNG = np.random.default_rng(2024)
# the reference's hard-coded derivative-test stencil and step (EDGE_EDGE_MOLLIFIER.h:385-438, 527-580)
REF_STENCIL = np.array([[0, 0, 0], [1, 0.1, 0], [0, 1.1, -0.1], [0, 0.1, -1.1]], np.float64)
REF_EPS = 1e-6
KINDS = {1: 3, 2: 4, 3: 4, 6: 4}  # oracle kind id -> number of points (PE, PT, EE, EE cross norm^2)


def fd_grad(f, x, eps):
    g = np.zeros(x.size)
    f0 = f(x)
    for i in range(x.size):
        y = x.copy().reshape(-1)
        y[i] += eps
        g[i] = (f(y.reshape(x.shape)) - f0) / eps
    return g


@pytest.mark.parametrize("kind", sorted(KINDS))
def test_closed_form_equals_autodiff(orc, kind):
    for _ in range(300):
        x = RNG.normal(size=(KINDS[kind], 3))
        g, H = orc.grad_hess(kind, x)
        gj, Hj = orc.grad_hess(kind, x, jet=True)
        assert np.abs(g - gj).max() <= 1e-11 * np.abs(gj).max()
        assert np.abs(H - Hj).max() <= 1e-11 * np.abs(Hj).max()
        assert np.abs(H - H.T).max() <= 1e-12 * np.abs(H).max()
        # translation invariance: gradient blocks sum to zero, Hessian annihilates translations
        assert np.abs(g.reshape(-1, 3).sum(0)).max() <= 1e-10 * np.abs(g).max()
        T = np.tile(np.eye(3), (KINDS[kind], 1))
        assert np.abs(H @ T).max() <= 1e-9 * np.abs(H).max()


@pytest.mark.parametrize("kind", [2, 3, 6])
def test_finite_differences_reference_stencil(orc, kind):
    """Same form as the reference's derivTest_*: forward differences with eps = 1e-6 on its hard-coded stencil."""
    x = REF_STENCIL
    g, H = orc.grad_hess(kind, x)
    fd = fd_grad(lambda y: orc.dist2(kind, y), x, REF_EPS)
    assert np.linalg.norm(g - fd) / np.linalg.norm(fd) < 1e-5
    Hfd = np.zeros((12, 12))
    for i in range(12):
        y = x.copy().reshape(-1)
        y[i] += REF_EPS
        Hfd[:, i] = (orc.grad_hess(kind, y.reshape(4, 3))[0] - g) / REF_EPS
    assert np.linalg.norm(H - Hfd) / np.linalg.norm(Hfd) < 1e-5


def test_mollifier_derivatives(orc):
    x = REF_STENCIL
    eps_x = 10.0  # derivTest_e default (EDGE_EDGE_MOLLIFIER.h:527)
    e, g, H = orc.mollifier(x, eps_x)
    fd = fd_grad(lambda y: orc.mollifier(y, eps_x)[0], x, REF_EPS)
    assert 0 < e < 1
    assert np.linalg.norm(g - fd) / np.linalg.norm(fd) < 1e-5
    Hfd = np.zeros((12, 12))
    for i in range(12):
        y = x.copy().reshape(-1)
        y[i] += REF_EPS
        Hfd[:, i] = (orc.mollifier(y.reshape(4, 3), eps_x)[1] - g) / REF_EPS
    assert np.linalg.norm(H - Hfd) / np.linalg.norm(Hfd) < 1e-5
    # beyond the threshold the mollifier is identically one
    e1, g1, H1 = orc.mollifier(x, 1e-9)
    assert e1 == 1.0 and not g1.any() and not H1.any()
    assert orc.mollifier_threshold(x) == 1e-3 * np.sum((x[0] - x[1]) ** 2) * np.sum((x[2] - x[3]) ** 2)


def test_barrier_scalars(orc):
    dhat2, kappa = 1e-4, 1e5
    for d in (1e-9, 3e-6, 5e-5, 8e-5):
        b, g, h = orc.barrier_scalar(d, dhat2, kappa)
        assert b > 0 and g < 0 and h > 0
        eps = d * 1e-6
        assert abs((orc.barrier_scalar(d + eps, dhat2, kappa)[0] - b) / eps - g) <= 1e-3 * abs(g)
        assert abs((orc.barrier_scalar(d + eps, dhat2, kappa)[1] - g) / eps - h) <= 1e-3 * abs(h)
    assert orc.barrier_scalar(dhat2, dhat2, kappa) == (0.0, 0.0, 0.0) or abs(orc.barrier_scalar(dhat2, dhat2, kappa)[0]) == 0.0


@pytest.mark.parametrize("n", [6, 9, 12])
def test_make_pd(orc, n):
    for _ in range(50):
        A = RNG.normal(size=(n, n))
        A = A + A.T
        P = orc.make_pd(A)
        lam, V = np.linalg.eigh(A)
        ref = (V * np.maximum(lam, 0)) @ V.T
        assert np.linalg.norm(P - ref) <= 1e-12 * np.linalg.norm(A)
        assert np.linalg.eigvalsh(P).min() >= -1e-12 * np.linalg.norm(A)
        assert np.linalg.norm(orc.make_pd(P) - P) <= 1e-12 * np.linalg.norm(A)  # idempotent
        S = A @ A.T + np.eye(n)
        assert np.array_equal(orc.make_pd(S), S)  # lambda_min >= 0: returned untouched (UTILS.h:13-15)
        # only the lower triangle is read (SelfAdjointEigenSolver)
        B = A.copy()
        B[np.triu_indices(n, 1)] = 123.0
        assert np.allclose(orc.make_pd(B), P, atol=1e-12 * np.linalg.norm(A))


def test_distance_types_cover_all_cases(orc):
    tri = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], np.float64)
    pts = {0: [-0.3, -0.3, 0.2], 1: [1.5, -0.2, 0.1], 2: [-0.2, 1.6, 0.1], 3: [0.5, -0.4, 0.3], 4: [0.8, 0.8, 0.2],
           5: [-0.5, 0.5, 0.1], 6: [0.2, 0.3, 0.4]}
    for t, p in pts.items():
        x = np.vstack([p, tri])
        assert orc.pt_type(x) == t
        # unclassified distance equals the brute-force distance to a dense sampling of the triangle
        u, v = np.meshgrid(np.linspace(0, 1, 201), np.linspace(0, 1, 201))
        m = u + v <= 1
        samples = tri[0] + u[m][:, None] * (tri[1] - tri[0]) + v[m][:, None] * (tri[2] - tri[0])
        brute = ((samples - np.array(p)) ** 2).sum(1).min()
        assert abs(orc.dist2(4, x) - brute) <= 1e-3 * brute + 1e-6
    ee = {8: ([0, 0, 0], [1, 0, 0], [0.5, -0.5, 0.3], [0.5, 0.5, 0.3]), 0: ([0, 0, 0], [1, 0, 0], [-1, -1, 0.5], [-2, -3, 1]),
          4: ([0, 0, 0], [1, 0, 0], [3, 3, 0.2], [1.5, 0.5, 0.1]), 6: ([0, 0, 0], [1, 0, 0], [0.5, 0.4, 0], [0.5, 2, 0.0])}
    for t, pts4 in ee.items():
        assert orc.ee_type(np.array(pts4, np.float64)) == t
    # parallel edges never classify as EE (case 8): the nearly-parallel guard of DISTANCE_TYPE.h:120
    assert orc.ee_type(np.array([[0, 0, 0], [1, 0, 0], [0.2, 0.1, 0], [0.8, 0.1, 0]], np.float64)) != 8


def test_hash_equals_brute_force(orc):
    """Ground truth for the broad phase: the reference's own all-pairs #else branches (IPC.h:166-168, 380-382)."""
    for name, m, d, dhats in make_cases():
        om = orc.mesh(m.X, m.X0, m.bnode, m.bedge, m.btri, m.dbc)
        for dh in dhats:
            for thickness in (0.0, 2e-3):
                a = orc.constraint_set(om, dh * dh, thickness, brute=False, want_cand=True)
                b = orc.constraint_set(om, dh * dh, thickness, brute=True, want_cand=True)
                assert np.array_equal(a[2], b[2]) and np.array_equal(a[3], b[3]), (name, dh)
                assert np.array_equal(lexsorted(a[0]), lexsorted(b[0])), (name, dh)
                assert np.array_equal(a[1], b[1])
        c1 = orc.ccd(om, d, 1.0, 0.0, brute=False, want_cand=True)
        c2 = orc.ccd(om, d, 1.0, 0.0, brute=True, want_cand=True)
        if c1["step_after_clamp"] == 1.0:  # no span clamp: same candidates and same step
            assert np.array_equal(c1["cand_pt"], c2["cand_pt"]) and np.array_equal(c1["cand_ee"], c2["cand_ee"]), name
            assert c1["step"] == c2["step"], name


def test_constraint_rows_encoding_and_merge(orc):
    name, m, d, dhats = make_cases()[0]
    om = orc.mesh(m.X, m.X0, m.bnode, m.bedge, m.btri, m.dbc)
    rows, info, cpt, cee = orc.constraint_set(om, dhats[-1] ** 2, want_cand=True)
    assert (rows[:, 1] >= 0).all()  # asserted by the reference in every decoder (IPC.h:803)
    merged = (rows[:, 0] < 0) & (rows[:, 3] < 0)
    mr = rows[merged]
    # merged group: key order of std::map<VECTOR<int,4>> over (k0,k1,k2), multiplicity >= 1 in slot 3, emitted last
    assert np.array_equal(mr[:, :3], lexsorted(mr[:, :3])) and (mr[:, 3] <= -1).all()
    assert merged[np.argmax(merged):].all()
    assert len(np.unique(mr[:, :3], axis=0)) == len(mr)
    assert (info[:, 0] == 1.0).all() and (info[:, 1] == dhats[-1] ** 2).all()  # OIPC weights (IPC.h:656-660)
    # every row lies below the activation distance; nothing below it is missed among the candidates' own minima
    d2, mn = orc.min_dist2(om, rows)
    assert (d2 < dhats[-1] ** 2).all() and mn == d2.min()


def test_energy_gradient_consistency(orc):
    name, m, d, dhats = make_cases()[1]
    dh, kappa = dhats[-1], 1e5
    om = orc.mesh(m.X, m.X0, m.bnode, m.bedge, m.btri, m.dbc)
    rows, info, _, _ = orc.constraint_set(om, dh * dh)
    st, E = orc.barrier(om, rows, info[:, 0], dh * dh, kappa)
    st, g = orc.barrier_gradient(om, rows, info[:, 0], dh * dh, kappa)
    assert st == 0 and E > 0
    idx = np.argsort(-np.abs(g).reshape(-1))[:6]
    for k in idx:  # directional finite differences on the six largest gradient entries, fixed constraint set
        X1 = m.X.copy().reshape(-1)
        h = 1e-9
        X1[k] += h
        om1 = orc.mesh(X1.reshape(-1, 3), m.X0, m.bnode, m.bedge, m.btri, m.dbc)
        E1 = orc.barrier(om1, rows, info[:, 0], dh * dh, kappa)[1]
        assert abs((E1 - E) / h - g.reshape(-1)[k]) <= 2e-3 * abs(g.reshape(-1)[k])
    # local row Hessians reproduce the assembled CSR (setFromTriplets semantics) and per-row E sums to E
    out = orc.barrier_hessian(om, rows, info[:, 0], dh * dh, kappa, project_spd=True, csr=True, triplets=True)
    import scipy.sparse as sp
    tr, tc, tv = out["triplets"]
    A = sp.coo_matrix((tv, (tr, tc)), shape=(3 * m.nV, 3 * m.nV)).tocsr()
    A.sum_duplicates(); A.sort_indices()
    ptr, col, val = out["csr"]
    # explicit zeros are kept: compare on the oracle's pattern
    B = sp.csr_matrix((val, col, ptr), shape=A.shape)
    assert abs(A - B).max() <= 1e-12 * abs(A).max()
    assert len(tv) == sum(144 if (r[0] >= 0 or r[3] >= 0) else (81 if r[2] >= 0 else 36) for r in rows)
    assert all(np.all(np.diff(col[ptr[i]:ptr[i + 1]]) > 0) for i in range(0, 3 * m.nV, 97))
    Esum = sum(orc.row_EgH(om, r, 1.0, dh * dh, kappa)[1] for r in rows[:: max(1, len(rows) // 400)])
    assert Esum > 0


def test_accd_is_conservative(orc):
    """Every hit returned by ACCD leaves a positive gap: distances at toc stay > 0 and the pair does not pass through."""
    rng = np.random.default_rng(5)
    hits = 0
    for _ in range(400):
        tri = rng.normal(size=(3, 3))
        n = np.cross(tri[1] - tri[0], tri[2] - tri[0]); n /= np.linalg.norm(n)
        bary = rng.dirichlet([1, 1, 1])
        p = bary @ tri + n * rng.uniform(0.05, 0.5)
        x = np.vstack([p, tri])
        dd = np.zeros((4, 3)); dd[0] = -n * rng.uniform(0.0, 1.5) + 0.1 * rng.normal(size=3)
        hit, toc, its = orc.accd(0, x, dd, 1.0)
        d0 = orc.dist2(4, x)
        if hit:
            hits += 1
            assert 0 < toc <= 1.0 and its >= 1
            d1 = orc.dist2(4, x + toc * dd)
            assert d1 > 0 and np.sqrt(d1) >= 0.05 * np.sqrt(d0)  # eta = 0.1: at least ~10% of the gap is kept
            # the point stays on its side of the triangle plane for every t <= toc
            side0 = np.dot(p - tri[0], n)
            for t in np.linspace(0, toc, 7):
                assert np.dot(p + t * dd[0] - tri[0], n) * side0 > 0
        else:
            assert toc == 1.0 or toc > 1.0 or toc == 1.0
    assert hits > 50
    # zero relative motion: no hit, toc untouched (CCD.h:297-299)
    x = np.vstack([[0.2, 0.2, 1.0], [0, 0, 0], [1, 0, 0], [0, 1, 0]])
    hit, toc, _ = orc.accd(0, x, np.tile([0.3, 0.1, -5.0], (4, 1)), 0.7)
    assert not hit and toc == 0.7


def test_ccd_step_is_intersection_free(orc):
    for name, m, d, dhats in make_cases()[:2]:
        om = orc.mesh(m.X, m.X0, m.bnode, m.bedge, m.btri, m.dbc)
        r = orc.ccd(om, 3.0 * d, 1.0)
        assert r["status"] == 0 and 0 < r["step"] <= 1.0
        X1 = m.X + r["step"] * 3.0 * d
        om1 = orc.mesh(X1, m.X0, m.bnode, m.bedge, m.btri, m.dbc)
        rows, _, _, _ = orc.constraint_set(om1, (2 * dhats[-1]) ** 2)
        if len(rows):
            d2, mn = orc.min_dist2(om1, rows)
            assert mn > 0
        # the span clamp (SPATIAL_HASH.h:477-482) caps the step before any pair is tested
        big = orc.ccd(om, 1e3 * d, 1.0)
        assert big["step_after_clamp"] < 1.0 and big["step"] <= big["step_after_clamp"]


def test_tree_mean_definition(orc):
    for n in (1, 2, 3, 7, 1000, 2049, 5000):
        a = RNG.uniform(0.5, 1.5, n)
        P = 1 << (n - 1).bit_length()
        b = np.zeros(max(P, 1)); b[:n] = a
        while len(b) > 1:
            b = b[0::2] + b[1::2]
        assert orc.tree_mean(a) == b[0] / n
