"""Synthetic codimensional surface meshes for the BASELINE.json configs, and the surface-primitive extraction
that fixes the input ordering contract of the hot path.

* `find_surface_primitives` follows Library/Utils/MESHIO.h:768-834 (Find_Surface_Primitives_And_Compute_Area):
  boundaryTri in element order, boundaryEdge = keys of a std::map<VECTOR<int,2>> (lexicographic), each undirected
  edge stored once with the orientation of the first triangle that introduced it, boundaryNode ascending.
* Generators restate SURVEY.md §8(d) configs 3-5 (seeded with numpy's PCG64 instead of mt19937_64; the seeds and
  parameters are the survey's).
"""
import numpy as np


# ------------------------------------------------------------------------------------------------
# primitive extraction (MESHIO.h:768-834)
# ------------------------------------------------------------------------------------------------
def find_surface_primitives(X, F):
    """X (nV,3) float64, F (nF,3) int -> dict(bnode, bedge, btri, BNArea, BEArea, BTArea)."""
    X = np.asarray(X, np.float64)
    F = np.ascontiguousarray(F, np.int32)
    nV = len(X)
    area = 0.5 * np.linalg.norm(np.cross(X[F[:, 1]] - X[F[:, 0]], X[F[:, 2]] - X[F[:, 0]]), axis=1)
    # directed edges in the reference's visiting order: (a,b), (b,c), (c,a) per triangle
    de = np.stack([F[:, [0, 1]], F[:, [1, 2]], F[:, [2, 0]]], axis=1).reshape(-1, 2).astype(np.int64)
    lo = np.minimum(de[:, 0], de[:, 1])
    hi = np.maximum(de[:, 0], de[:, 1])
    key = lo * nV + hi
    _, first, inv = np.unique(key, return_index=True, return_inverse=True)
    stored = de[first]  # orientation of the first directed edge that introduced the undirected edge
    earea = np.bincount(inv, weights=np.repeat(area / 3.0, 3), minlength=len(first))
    order = np.lexsort((stored[:, 1], stored[:, 0]))
    bedge = stored[order].astype(np.int32)
    BEArea = earea[order] / 2.0
    narea = np.bincount(F.reshape(-1), weights=np.repeat(area / 3.0, 3), minlength=nV)
    bnode = np.nonzero(narea != 0)[0].astype(np.int32)
    return {
        "bnode": bnode, "bedge": np.ascontiguousarray(bedge), "btri": F.copy(),
        "BNArea": narea[bnode], "BEArea": BEArea, "BTArea": area / 2.0,
    }


class SurfaceMesh:
    """Positions + triangles + the ordered surface primitives the hot path receives."""

    def __init__(self, X, F, X0=None, dbc=None):
        self.X = np.ascontiguousarray(X, np.float64)
        self.X0 = self.X.copy() if X0 is None else np.ascontiguousarray(X0, np.float64)
        self.F = np.ascontiguousarray(F, np.int32)
        p = find_surface_primitives(self.X0, self.F)
        self.bnode, self.bedge, self.btri = p["bnode"], p["bedge"], p["btri"]
        self.areas = (p["BNArea"], p["BEArea"], p["BTArea"])
        self.dbc = np.zeros(len(self.X), np.uint8) if dbc is None else np.ascontiguousarray(dbc, np.uint8)

    @property
    def nV(self):
        return len(self.X)

    @property
    def nF(self):
        return len(self.F)

    def vertex_normals(self):
        n = np.cross(self.X[self.F[:, 1]] - self.X[self.F[:, 0]], self.X[self.F[:, 2]] - self.X[self.F[:, 0]])
        vn = np.zeros_like(self.X)
        for k in range(3):
            np.add.at(vn, self.F[:, k], n)
        ln = np.linalg.norm(vn, axis=1, keepdims=True)
        return vn / np.where(ln > 0, ln, 1.0)


# ------------------------------------------------------------------------------------------------
# generators
# ------------------------------------------------------------------------------------------------
def icosphere(nu, radius=1.0):
    """Class-I geodesic icosphere of frequency nu: F = 20 nu^2, V = 10 nu^2 + 2."""
    t = (1.0 + 5.0 ** 0.5) / 2.0
    P = np.array([[-1, t, 0], [1, t, 0], [-1, -t, 0], [1, -t, 0], [0, -1, t], [0, 1, t], [0, -1, -t], [0, 1, -t],
                  [t, 0, -1], [t, 0, 1], [-t, 0, -1], [-t, 0, 1]], np.float64)
    T = np.array([[0, 11, 5], [0, 5, 1], [0, 1, 7], [0, 7, 10], [0, 10, 11], [1, 5, 9], [5, 11, 4], [11, 10, 2],
                  [10, 7, 6], [7, 1, 8], [3, 9, 4], [3, 4, 2], [3, 2, 6], [3, 6, 8], [3, 8, 9], [4, 9, 5], [2, 4, 11],
                  [6, 2, 10], [8, 6, 7], [9, 8, 1]], np.int64)
    ii, jj = np.meshgrid(np.arange(nu + 1), np.arange(nu + 1), indexing="ij")
    mask = ii + jj <= nu
    ii, jj = ii[mask], jj[mask]
    kk = nu - ii - jj
    lid = -np.ones((nu + 1, nu + 1), np.int64)
    lid[ii, jj] = np.arange(len(ii))
    # small triangles of one face in local ids
    a, b = np.meshgrid(np.arange(nu), np.arange(nu), indexing="ij")
    up = (a + b) <= nu - 1
    tri_up = np.stack([lid[a[up], b[up]], lid[a[up] + 1, b[up]], lid[a[up], b[up] + 1]], axis=1)
    dn = (a + b) <= nu - 2
    tri_dn = np.stack([lid[a[dn] + 1, b[dn]], lid[a[dn] + 1, b[dn] + 1], lid[a[dn], b[dn] + 1]], axis=1)
    local_tris = np.concatenate([tri_up, tri_dn], axis=0)
    pts, tris = [], []
    for f, (A, B, Cc) in enumerate(T):
        pts.append((kk[:, None] * P[A] + ii[:, None] * P[B] + jj[:, None] * P[Cc]) / nu)
        tris.append(local_tris + f * len(ii))
    pts = np.concatenate(pts, axis=0)
    tris = np.concatenate(tris, axis=0)
    q = np.round(pts * 1e7).astype(np.int64)
    _, first, inv = np.unique(q, axis=0, return_index=True, return_inverse=True)
    inv = inv.reshape(-1)
    V = pts[first]
    V = radius * V / np.linalg.norm(V, axis=1, keepdims=True)
    Fm = inv[tris].astype(np.int32)
    return V, Fm


def nested_icospheres(nu=112, gap=5e-3, jitter=1e-4, seed=20260101):
    """Config 3: two concentric icospheres (radii 1 and 1+gap), jitter along the normal; returns (mesh, searchDir)."""
    rng = np.random.default_rng(seed)
    V1, F1 = icosphere(nu, 1.0)
    V2, F2 = icosphere(nu, 1.0 + gap)
    n1 = V1 / np.linalg.norm(V1, axis=1, keepdims=True)
    n2 = V2 / np.linalg.norm(V2, axis=1, keepdims=True)
    V1 = V1 + n1 * rng.uniform(-jitter, jitter, (len(V1), 1))
    V2 = V2 + n2 * rng.uniform(-jitter, jitter, (len(V2), 1))
    X = np.concatenate([V1, V2], axis=0)
    F = np.concatenate([F1, F2 + len(V1)], axis=0)
    s = 0.6 * gap
    direction = np.concatenate([n1 * s, -n2 * s], axis=0)
    return SurfaceMesh(X, F), np.ascontiguousarray(direction)


def sheet_stack(n_sheets=8, nx=500, ny=500, h=4e-3, A=1.5e-3, jitter=1e-5, seed=20260102, dir_sigma=1e-3,
                dir_seed=20260103, extent=(1.0, 1.0)):
    """Configs 4/5: n_sheets wavy sheets of (nx+1)x(ny+1) vertices, 2*nx*ny triangles each, stacked h apart.

    Sheet k sits at z = k*h + A sin(2 pi (fx x + px)) sin(2 pi (fy y + py)); neighbouring sheets never intersect
    because 2A < h. Returns (mesh, searchDir) with searchDir = N(0, dir_sigma^2) + a per-sheet z-approach term.
    """
    rng = np.random.default_rng(seed)
    gx = np.linspace(0.0, extent[0], nx + 1)
    gy = np.linspace(0.0, extent[1], ny + 1)
    xx, yy = np.meshgrid(gx, gy, indexing="ij")
    idx = np.arange((nx + 1) * (ny + 1)).reshape(nx + 1, ny + 1)
    t1 = np.stack([idx[:-1, :-1], idx[1:, :-1], idx[:-1, 1:]], axis=-1).reshape(-1, 3)
    t2 = np.stack([idx[1:, 1:], idx[:-1, 1:], idx[1:, :-1]], axis=-1).reshape(-1, 3)
    Fs = np.concatenate([t1, t2], axis=0)
    nvs = (nx + 1) * (ny + 1)
    Xs, Fl, Ds = [], [], []
    jr = np.random.default_rng(seed + 1)
    dr = np.random.default_rng(dir_seed)
    for k in range(n_sheets):
        fx, fy = rng.integers(3, 10, 2)
        px, py = rng.uniform(0, 1, 2)
        z = k * h + A * np.sin(2 * np.pi * (fx * xx + px)) * np.sin(2 * np.pi * (fy * yy + py))
        P = np.stack([xx, yy, z], axis=-1).reshape(-1, 3)
        P = P + jr.uniform(-jitter, jitter, P.shape)
        Xs.append(P)
        Fl.append(Fs + k * nvs)
        d = dr.normal(0.0, dir_sigma, P.shape)
        d[:, 2] += (-1.0 if k % 2 else 1.0) * 0.25 * h  # neighbouring sheets approach each other
        Ds.append(d)
    X = np.concatenate(Xs, axis=0)
    F = np.concatenate(Fl, axis=0).astype(np.int32)
    return SurfaceMesh(X, F), np.ascontiguousarray(np.concatenate(Ds, axis=0))


def random_soup(n_tris=200, seed=0, scale=1.0, tri_size=0.15):
    """Small random triangle soup (disconnected triangles) for brute-force parity tests."""
    rng = np.random.default_rng(seed)
    c = rng.uniform(0, scale, (n_tris, 1, 3))
    X = (c + rng.uniform(-tri_size, tri_size, (n_tris, 3, 3))).reshape(-1, 3)
    F = np.arange(3 * n_tris, dtype=np.int32).reshape(-1, 3)
    return SurfaceMesh(X, F)
