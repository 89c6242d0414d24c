"""ORACLE — TEST INFRASTRUCTURE ONLY. ctypes view of oracle/_ref/libidp_ref.so: the reference's own per-pair math
(Library/Math/Distance/*.h, BARRIER.h, UTILS.h) compiled from /root/reference by oracle/ref_shim/Makefile."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(_HERE, "_ref", "libidp_ref.so")
LIB_IPC = os.path.join(_HERE, "_ref", "libidp_ref_ipc.so")
LIB_CSR = os.path.join(_HERE, "_ref", "libidp_ref_csr.so")


def available():
    return os.path.exists(LIB)


def ipc_available():
    return os.path.exists(LIB_IPC)


def csr_available():
    return os.path.exists(LIB_CSR)


def ref_csr_system(n, tr, tc, tv, mdiag=None, dbc=None, dim=3):
    """The reference's own Math/CSR_MATRIX.h (oracle/ref_shim/ref_csr_capi.cpp): Construct_From_Triplet -> += M ->
    Project_DBC, the sequence of FEM/Shell/INC_POTENTIAL.h:382-394. mdiag: n scalars (0 = no entry), dbc: n/dim flags."""
    L = C.CDLL(LIB_CSR)
    L.ref_csr_system.restype = C.c_long
    L.ref_csr_system.argtypes = [C.c_int, C.c_long] + [C.c_void_p] * 5 + [C.c_int] + [C.c_void_p] * 3 + [C.c_long]
    tr = np.ascontiguousarray(tr, np.int32); tc = np.ascontiguousarray(tc, np.int32); tv = np.ascontiguousarray(tv, np.float64)
    mdiag = None if mdiag is None else np.ascontiguousarray(mdiag, np.float64)
    dbc = None if dbc is None else np.ascontiguousarray(dbc, np.uint8)
    cap = len(tv) + n
    ptr = np.empty(n + 1, np.int32); col = np.empty(cap, np.int32); val = np.empty(cap, np.float64)
    nnz = L.ref_csr_system(n, len(tv), _p(tr), _p(tc), _p(tv), None if mdiag is None else _p(mdiag), None if dbc is None else _p(dbc), dim,
                           _p(ptr), _p(col), _p(val), cap)
    assert nnz >= 0
    return ptr, col[:nnz].copy(), val[:nnz].copy()


def build():
    """Only possible where /root/reference is mounted (the authoring container)."""
    if os.path.isdir("/root/reference/Library"):
        subprocess.check_call(["make", "-C", os.path.join(_HERE, "ref_shim")], stdout=subprocess.DEVNULL)
    return available()


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class Reference:
    def __init__(self):
        self.lib = C.CDLL(LIB)
        L = self.lib
        L.ref_dist2.restype = C.c_double
        L.ref_dist2.argtypes = [C.c_int, C.c_void_p]
        L.ref_grad_hess.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ref_mollifier.argtypes = [C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ref_mollifier_threshold.restype = C.c_double
        L.ref_barrier_scalar.argtypes = [C.c_double, C.c_double, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ref_accd.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_void_p]
        L.ref_aabb.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_double]

    def pt_type(self, x):
        x = np.ascontiguousarray(x, np.float64); return self.lib.ref_pt_type(_p(x))

    def ee_type(self, x):
        x = np.ascontiguousarray(x, np.float64); return self.lib.ref_ee_type(_p(x))

    def dist2(self, kind, x):
        x = np.ascontiguousarray(x, np.float64); return self.lib.ref_dist2(kind, _p(x))

    def grad_hess(self, kind, x):
        n = {0: 6, 1: 9, 2: 12, 3: 12, 6: 12}[kind]
        x = np.ascontiguousarray(x, np.float64); g = np.zeros(n); H = np.zeros((n, n))
        self.lib.ref_grad_hess(kind, _p(x), _p(g), _p(H))
        return g, H.T.copy()  # column-major -> row-major

    def mollifier(self, x, eps_x):
        x = np.ascontiguousarray(x, np.float64); e = C.c_double(0); g = np.zeros(12); H = np.zeros((12, 12))
        self.lib.ref_mollifier(_p(x), eps_x, C.byref(e), _p(g), _p(H))
        return e.value, g, H.T.copy()

    def mollifier_threshold(self, x0):
        x0 = np.ascontiguousarray(x0, np.float64); return self.lib.ref_mollifier_threshold(_p(x0))

    def barrier_scalar(self, d, dhat2, kappa):
        b = C.c_double(0); g = C.c_double(0); h = C.c_double(0)
        self.lib.ref_barrier_scalar(d, dhat2, kappa, C.byref(b), C.byref(g), C.byref(h))
        return b.value, g.value, h.value

    def make_pd12(self, H):
        M = np.ascontiguousarray(np.asarray(H, np.float64).T).copy()
        self.lib.ref_make_pd12(_p(M)); return M.T.copy()

    def accd(self, kind, x, d, toc, eta=0.1, thickness=0.0):
        x = np.ascontiguousarray(x, np.float64); d = np.ascontiguousarray(d, np.float64); t = C.c_double(toc)
        hit = self.lib.ref_accd(kind, _p(x), _p(d), eta, thickness, C.byref(t))
        return bool(hit), t.value

    def aabb(self, kind, x, d, dist):
        x = np.ascontiguousarray(x, np.float64)
        d = np.ascontiguousarray(d if d is not None else np.zeros_like(x), np.float64)
        return bool(self.lib.ref_aabb(kind, _p(x), _p(d), dist))


class ReferenceIPC:
    """The reference's own six contact operators (FEM/IPC.h + Grid/SPATIAL_HASH.h compiled from /root/reference on
    std::vector storage stand-ins, oracle/ref_shim/ref_ipc_capi.cpp): <double, 3, shell=false, elasticIPC=false>."""

    def __init__(self):
        L = self.lib = C.CDLL(LIB_IPC)
        L.refipc_constraint_set.restype = C.c_int
        L.refipc_constraint_set.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p,
                                            C.c_void_p, C.c_double, C.c_double, C.c_long, C.c_void_p, C.c_void_p]
        L.refipc_barrier.restype = C.c_long
        L.refipc_barrier.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_double,
                                     C.c_int, C.c_void_p, C.c_void_p, C.c_long, C.c_void_p, C.c_void_p, C.c_void_p]
        L.refipc_ccd.restype = C.c_double
        L.refipc_ccd.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                 C.c_double, C.c_double]
        L.refipc_min_dist2.restype = C.c_double
        L.refipc_min_dist2.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_double, C.c_void_p]
        if hasattr(L, "refipc_friction"):
            L.refipc_friction.restype = C.c_long
            L.refipc_friction.argtypes = ([C.c_int] + [C.c_void_p] * 3 + [C.c_int, C.c_void_p, C.c_void_p] + [C.c_double] * 5 + [C.c_int] +
                                          [C.c_void_p] * 7 + [C.c_long] + [C.c_void_p] * 3)

    @staticmethod
    def _mesh(m):
        return (np.ascontiguousarray(m.X, np.float64), np.ascontiguousarray(m.X0, np.float64), np.ascontiguousarray(m.bnode, np.int32),
                np.ascontiguousarray(m.bedge, np.int32), np.ascontiguousarray(m.btri, np.int32), np.ascontiguousarray(m.dbc, np.uint8))

    def constraint_set(self, m, dHat2, thickness=0.0, cap=4000000):
        X, X0, bn, be, bt, dbc = self._mesh(m)
        rows = np.zeros((cap, 4), np.int32); info = np.zeros((cap, 2))
        n = self.lib.refipc_constraint_set(len(X), _p(X), _p(X0), len(bn), _p(bn), len(be), _p(be), len(bt), _p(bt), _p(dbc), dHat2, thickness,
                                           cap, _p(rows), _p(info))
        assert n <= cap
        return rows[:n].copy(), info[:n].copy()

    def barrier(self, m, rows, weights, dHat2, kappa, thickness=0.0, project_spd=True, want_h=True):
        X, X0 = np.ascontiguousarray(m.X, np.float64), np.ascontiguousarray(m.X0, np.float64)
        rows = np.ascontiguousarray(rows, np.int32); w = np.ascontiguousarray(weights, np.float64)
        E = C.c_double(0.0); g = np.zeros((len(X), 3))
        cap = 144 * len(rows) + 1 if want_h else 1
        tr = np.zeros(cap, np.int32); tc = np.zeros(cap, np.int32); tv = np.zeros(cap)
        nt = self.lib.refipc_barrier(len(X), _p(X), _p(X0), len(rows), _p(rows), _p(w), dHat2, kappa, thickness, int(project_spd), C.byref(E), _p(g),
                                     cap, _p(tr) if want_h else None, _p(tc) if want_h else None, _p(tv) if want_h else None)
        return E.value, g, (tr[:nt], tc[:nt], tv[:nt])

    def ccd(self, m, direction, step=1.0, thickness=0.0):
        X, X0, bn, be, bt, dbc = self._mesh(m)
        d = np.ascontiguousarray(direction, np.float64)
        return self.lib.refipc_ccd(len(X), _p(X), len(bn), _p(bn), len(be), _p(be), len(bt), _p(bt), _p(dbc), _p(d), thickness, step)

    def min_dist2(self, m, rows, thickness=0.0):
        X = np.ascontiguousarray(m.X, np.float64); rows = np.ascontiguousarray(rows, np.int32)
        d = np.zeros(len(rows))
        mn = self.lib.refipc_min_dist2(len(X), _p(X), len(rows), _p(rows), thickness, _p(d))
        return d, mn

    def friction(self, Xb, rows, dHat2, kappa, X=None, Xn=None, epsv2_h2=1e-6, mu=0.3, thickness=0.0, project_spd=True, weights=None):
        """The reference's own FEM/FRICTION.h: Compute_Friction_Basis at Xb (returned as rows / closest points / bases / normal
        forces) and, with X and Xn, potential, gradient and Hessian triplets."""
        Xb = np.ascontiguousarray(Xb, np.float64); rows = np.ascontiguousarray(rows, np.int32)
        n = len(rows)
        w = np.ones(n) if weights is None else np.ascontiguousarray(weights, np.float64)
        nf = C.c_int(0)
        frows = np.zeros((n, 4), np.int32); cp = np.zeros((n, 2)); basis = np.zeros((n, 6)); lam = np.zeros(n)
        E = C.c_double(0.0); g = np.zeros_like(Xb)
        ev = X is not None
        if ev:
            X = np.ascontiguousarray(X, np.float64); Xn = np.ascontiguousarray(Xn, np.float64)
        cap = 144 * n + 1
        tr = np.zeros(cap, np.int32); tc = np.zeros(cap, np.int32); tv = np.zeros(cap)
        nt = self.lib.refipc_friction(len(Xb), _p(Xb), _p(X) if ev else None, _p(Xn) if ev else None, n, _p(rows), _p(w), dHat2, kappa, thickness,
                                      epsv2_h2, mu, int(project_spd), C.byref(nf), _p(frows), _p(cp), _p(basis), _p(lam), C.byref(E), _p(g), cap,
                                      _p(tr), _p(tc), _p(tv))
        k = nf.value
        return dict(rows=frows[:k], closest=cp[:k], basis=basis[:k], normal_force=lam[:k], E=E.value, g=g, triplets=(tr[:nt], tc[:nt], tv[:nt]))
