"""ORACLE — TEST INFRASTRUCTURE ONLY.

ctypes binding of the CPU restatement in oracle/ (capi.cpp). Only tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py may import this module; nothing under idp_b200/ does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_BUILD = os.path.join(_HERE, "_build")


def build(force=False):
    """Compile the two oracle builds (parity: -ffp-contract=off; fast: the reference's -O3 -mfma flags)."""
    outs = [os.path.join(_BUILD, n) for n in ("liborc_parity.so", "liborc_fast.so")]
    srcs = [os.path.join(_HERE, n) for n in ("capi.cpp", "orc_math.hpp", "orc_deriv.hpp", "orc_ipc.hpp", "orc_system.hpp", "Makefile")]
    stale = force or any(not os.path.exists(o) for o in outs) or \
        max(os.path.getmtime(s) for s in srcs) > min(os.path.getmtime(o) for o in outs)
    if stale:
        subprocess.check_call(["make", "-C", _HERE, "-j2"], stdout=subprocess.DEVNULL)
    return outs


class OrcMesh(C.Structure):
    _fields_ = [("nV", C.c_int), ("X", C.c_void_p), ("X0", C.c_void_p),
                ("nBN", C.c_int), ("bnode", C.c_void_p), ("nBE", C.c_int), ("bedge", C.c_void_p),
                ("nBT", C.c_int), ("btri", C.c_void_p), ("dbc", C.c_void_p)]


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Oracle:
    """One loaded oracle build. `flavor` is "parity" (default) or "fast"."""

    def __init__(self, flavor="parity"):
        build()
        self.lib = C.CDLL(os.path.join(_BUILD, "liborc_%s.so" % flavor))
        L = self.lib
        L.orc_constraint_set.restype = C.c_void_p
        L.orc_constraint_set.argtypes = [C.POINTER(OrcMesh), C.c_double, C.c_double, C.c_int, C.c_int]
        L.orc_cs_count.restype = C.c_long
        L.orc_cs_count.argtypes = [C.c_void_p, C.c_int]
        L.orc_cs_copy.argtypes = [C.c_void_p] * 5
        L.orc_cs_free.argtypes = [C.c_void_p]
        for f in (L.orc_barrier, L.orc_barrier_gradient):
            f.restype = C.c_int
            f.argtypes = [C.POINTER(OrcMesh), C.c_long, C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_void_p]
        L.orc_barrier_hessian.restype = C.c_void_p
        L.orc_barrier_hessian.argtypes = [C.POINTER(OrcMesh), C.c_long, C.c_void_p, C.c_void_p, C.c_double, C.c_double,
                                          C.c_double, C.c_int, C.c_int]
        L.orc_hess_status.argtypes = [C.c_void_p]
        L.orc_hess_count.restype = C.c_long
        L.orc_hess_count.argtypes = [C.c_void_p, C.c_int]
        L.orc_hess_copy.argtypes = [C.c_void_p] * 7
        L.orc_hess_free.argtypes = [C.c_void_p]
        L.orc_row_EgH.restype = C.c_int
        L.orc_row_EgH.argtypes = [C.POINTER(OrcMesh), C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int,
                                  C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_min_dist2.argtypes = [C.POINTER(OrcMesh), C.c_long, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p]
        L.orc_ccd.restype = C.c_void_p
        L.orc_ccd.argtypes = [C.POINTER(OrcMesh), C.c_void_p, C.c_double, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                              C.c_void_p, C.c_void_p]
        L.orc_ccd_count.restype = C.c_long
        L.orc_ccd_count.argtypes = [C.c_void_p, C.c_int]
        L.orc_ccd_copy.argtypes = [C.c_void_p] * 3
        L.orc_ccd_free.argtypes = [C.c_void_p]
        L.orc_pt_type.argtypes = [C.c_void_p]
        L.orc_ee_type.argtypes = [C.c_void_p]
        L.orc_dist2.restype = C.c_double
        L.orc_dist2.argtypes = [C.c_int, C.c_void_p]
        L.orc_grad_hess.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_mollifier.argtypes = [C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_mollifier_threshold.restype = C.c_double
        L.orc_mollifier_threshold.argtypes = [C.c_void_p]
        L.orc_barrier_scalar.argtypes = [C.c_double, C.c_double, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_make_pd.argtypes = [C.c_int, C.c_void_p]
        L.orc_sym_eig.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_accd.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_void_p, C.c_void_p]
        L.orc_aabb.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_double]
        L.orc_tree_mean.restype = C.c_double
        L.orc_tree_mean.argtypes = [C.c_void_p, C.c_long]
        L.orc_mean_edge_length.restype = C.c_double
        L.orc_mean_edge_length.argtypes = [C.POINTER(OrcMesh)]

    # ---- mesh marshalling ----
    @staticmethod
    def mesh(X, X0, bnode, bedge, btri, dbc=None):
        """Returns (OrcMesh, keepalive tuple). X/X0: (nV,3) float64; bnode (nBN,), bedge (nBE,2), btri (nBT,3) int32."""
        X = _f64(X); X0 = _f64(X0 if X0 is not None else X)
        bnode = _i32(bnode); bedge = _i32(bedge); btri = _i32(btri)
        dbc = np.ascontiguousarray(dbc if dbc is not None else np.zeros(len(X), np.uint8), dtype=np.uint8)
        m = OrcMesh(len(X), _p(X), _p(X0), len(bnode), _p(bnode), len(bedge), _p(bedge), len(btri), _p(btri), _p(dbc))
        return m, (X, X0, bnode, bedge, btri, dbc)

    # ---- the six operators ----
    def constraint_set(self, mesh, dHat2, thickness=0.0, brute=False, want_cand=False):
        m, _keep = mesh
        h = self.lib.orc_constraint_set(C.byref(m), dHat2, thickness, int(brute), int(want_cand))
        n = self.lib.orc_cs_count(h, 0)
        rows = np.empty((n, 4), np.int32); info = np.empty((n, 2), np.float64)
        cpt = np.empty((self.lib.orc_cs_count(h, 1), 2), np.int32)
        cee = np.empty((self.lib.orc_cs_count(h, 2), 2), np.int32)
        self.lib.orc_cs_copy(h, _p(rows), _p(info), _p(cpt), _p(cee))
        self.lib.orc_cs_free(h)
        return rows, info, cpt, cee

    def barrier(self, mesh, rows, weight, dHat2, kappa, thickness=0.0):
        m, _keep = mesh
        rows = _i32(rows); weight = _f64(weight)
        E = C.c_double(0.0)
        st = self.lib.orc_barrier(C.byref(m), len(rows), _p(rows), _p(weight), dHat2, kappa, thickness, C.byref(E))
        return st, E.value

    def barrier_gradient(self, mesh, rows, weight, dHat2, kappa, thickness=0.0):
        m, keep = mesh
        rows = _i32(rows); weight = _f64(weight)
        g = np.zeros((m.nV, 3), np.float64)
        st = self.lib.orc_barrier_gradient(C.byref(m), len(rows), _p(rows), _p(weight), dHat2, kappa, thickness, _p(g))
        return st, g

    def barrier_hessian(self, mesh, rows, weight, dHat2, kappa, thickness=0.0, project_spd=True, csr=True, triplets=False):
        m, _keep = mesh
        rows = _i32(rows); weight = _f64(weight)
        h = self.lib.orc_barrier_hessian(C.byref(m), len(rows), _p(rows), _p(weight), dHat2, kappa, thickness,
                                         int(project_spd), int(csr))
        st = self.lib.orc_hess_status(h)
        out = {"status": st}
        if triplets:
            nt = self.lib.orc_hess_count(h, 0)
            tr = np.empty(nt, np.int32); tc = np.empty(nt, np.int32); tv = np.empty(nt, np.float64)
            self.lib.orc_hess_copy(h, _p(tr), _p(tc), _p(tv), None, None, None)
            out["triplets"] = (tr, tc, tv)
        if csr and st == 0:
            nnz = self.lib.orc_hess_count(h, 1)
            ptr = np.empty(3 * m.nV + 1, np.int32); col = np.empty(nnz, np.int32); val = np.empty(nnz, np.float64)
            self.lib.orc_hess_copy(h, None, None, None, _p(ptr), _p(col), _p(val))
            out["csr"] = (ptr, col, val)
        self.lib.orc_hess_free(h)
        return out

    def system_matrix(self, mesh, rows, weight, dHat2, kappa, thickness=0.0, project_spd=True, elem=None, vol=None, h=0.0, mass=None,
                      project_dbc=False, want_triplets=False):
        """[flow triplets][barrier triplets] -> CSR -> += M -> Project_DBC (orc_system.hpp); returns (ptr, col, val)."""
        L = self.lib
        L.orc_system_matrix.restype = C.c_void_p
        L.orc_system_matrix.argtypes = [C.POINTER(OrcMesh), C.c_long, C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_int,
                                        C.c_int, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_int]
        m, _keep = mesh
        rows = _i32(rows).reshape(-1, 4)
        weight = _f64(weight)
        nE = 0 if elem is None else len(elem)
        elem = None if elem is None else _i32(elem)
        vol = None if vol is None else _f64(vol)
        mass = None if mass is None else _f64(mass)
        hnd = L.orc_system_matrix(C.byref(m), len(rows), _p(rows), _p(weight), dHat2, kappa, thickness, int(project_spd), nE, _p(elem), _p(vol), h,
                                  _p(mass), int(project_dbc))
        try:
            if L.orc_hess_status(hnd) != 0:
                raise RuntimeError("oracle system matrix: status %d" % L.orc_hess_status(hnd))
            nnz = L.orc_hess_count(hnd, 1)
            ptr = np.empty(3 * m.nV + 1, np.int32); col = np.empty(nnz, np.int32); val = np.empty(nnz, np.float64)
            L.orc_hess_copy(hnd, None, None, None, _p(ptr), _p(col), _p(val))
            if want_triplets:
                nt = L.orc_hess_count(hnd, 0)
                tr = np.empty(nt, np.int32); tc = np.empty(nt, np.int32); tv = np.empty(nt, np.float64)
                L.orc_hess_copy(hnd, _p(tr), _p(tc), _p(tv), None, None, None)
        finally:
            L.orc_hess_free(hnd)
        if want_triplets:
            return ptr, col, val, (tr, tc, tv)
        return ptr, col, val

    def surface(self, nV, tri, X):
        """Find_Surface_Primitives_And_Compute_Area restated with std::map (orc_system.hpp)."""
        L = self.lib
        L.orc_surface.restype = C.c_void_p
        L.orc_surface.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_surface_count.restype = C.c_long
        L.orc_surface_count.argtypes = [C.c_void_p, C.c_int]
        L.orc_surface_copy.argtypes = [C.c_void_p] * 7
        L.orc_surface_free.argtypes = [C.c_void_p]
        tri = _i32(tri).reshape(-1, 3)
        X = _f64(X).reshape(-1, 3)
        hnd = L.orc_surface(nV, len(tri), _p(tri), _p(X))
        nN, nE, nT = (L.orc_surface_count(hnd, k) for k in range(3))
        out = dict(bnode=np.empty(nN, np.int32), bedge=np.empty((nE, 2), np.int32), btri=np.empty((nT, 3), np.int32),
                   BNArea=np.empty(nN), BEArea=np.empty(nE), BTArea=np.empty(nT))
        L.orc_surface_copy(hnd, _p(out["bnode"]), _p(out["bedge"]), _p(out["btri"]), _p(out["BNArea"]), _p(out["BEArea"]), _p(out["BTArea"]))
        L.orc_surface_free(hnd)
        return out

    def membrane_batch(self, X, elem, ib, coef, lam, mu, dbc=None, project_spd=True, want_h=True):
        """Membrane E / g / H per triangle (orc_elastic.hpp; MEMBRANE.h:8-315): returns E[e], g (nV x 3), H (nE x 9 x 9), active[e]."""
        L = self.lib
        L.orc_membrane_batch.argtypes = [C.c_int] + [C.c_void_p] * 7 + [C.c_int] + [C.c_void_p] * 4
        X = _f64(X).reshape(-1, 3); elem = _i32(elem).reshape(-1, 3); ib = _f64(ib).reshape(-1, 3)
        n = len(elem)
        coef, lam, mu = (_f64(np.broadcast_to(a, n)) for a in (coef, lam, mu))
        dbc = None if dbc is None else np.ascontiguousarray(dbc, np.uint8)
        E = np.zeros(n); g = np.zeros_like(X); H = np.zeros((n, 9, 9)) if want_h else None; act = np.zeros(n, np.uint8)
        L.orc_membrane_batch(n, _p(elem), _p(X), _p(ib), _p(coef), _p(lam), _p(mu), None if dbc is None else _p(dbc), int(project_spd), _p(E), _p(g),
                             None if H is None else _p(H), _p(act))
        return E, g, H, act

    def hinge_batch(self, X, stencil, info, kh2, dbc=None, project_spd=True, want_h=True):
        """Hinge bending E / g / H (orc_elastic.hpp; BENDING.h KL=false): info = (thetabar, ebar, hbar) per hinge, kh2 = h^2 k."""
        L = self.lib
        L.orc_hinge_batch.argtypes = [C.c_int] + [C.c_void_p] * 3 + [C.c_double, C.c_void_p, C.c_int] + [C.c_void_p] * 4
        X = _f64(X).reshape(-1, 3); stencil = _i32(stencil).reshape(-1, 4); info = _f64(info).reshape(-1, 3)
        n = len(stencil)
        dbc = None if dbc is None else np.ascontiguousarray(dbc, np.uint8)
        E = np.zeros(n); g = np.zeros_like(X); H = np.zeros((n, 12, 12)) if want_h else None; act = np.zeros(n, np.uint8)
        L.orc_hinge_batch(n, _p(stencil), _p(X), _p(info), float(kh2), None if dbc is None else _p(dbc), int(project_spd), _p(E), _p(g),
                          None if H is None else _p(H), _p(act))
        return E, g, H, act

    def dihedral_angle(self, x12):
        self.lib.orc_dihedral_angle.restype = C.c_double
        self.lib.orc_dihedral_angle.argtypes = [C.c_void_p]
        x12 = _f64(x12)
        return self.lib.orc_dihedral_angle(_p(x12))

    def friction(self, Xb, rows, dHat2, kappa, X=None, Xn=None, epsv2_h2=1e-6, mu=0.3, thickness=0.0, project_spd=True, weights=None):
        """Lagged friction (orc_friction.hpp; FEM/FRICTION.h): basis at Xb, then E / g / triplets at X relative to Xn. Same
        return layout as oracle.ref_binding.ReferenceIPC.friction."""
        L = self.lib
        L.orc_friction.restype = C.c_long
        L.orc_friction.argtypes = ([C.c_int] + [C.c_void_p] * 3 + [C.c_long, C.c_void_p, C.c_void_p] + [C.c_double] * 5 + [C.c_int] + [C.c_void_p] * 7 +
                                   [C.c_long] + [C.c_void_p] * 3)
        Xb = _f64(Xb).reshape(-1, 3); rows = _i32(rows).reshape(-1, 4)
        n = len(rows)
        w = np.ones(n) if weights is None else _f64(weights)
        nf = C.c_int(0)
        frows = np.zeros((n, 4), np.int32); cp = np.zeros((n, 2)); basis = np.zeros((n, 6)); lam = np.zeros(n)
        E = C.c_double(0.0); g = np.zeros_like(Xb)
        ev = X is not None
        if ev:
            X = _f64(X).reshape(-1, 3); Xn = _f64(Xn).reshape(-1, 3)
        cap = 144 * n + 1
        tr = np.zeros(cap, np.int32); tc = np.zeros(cap, np.int32); tv = np.zeros(cap)
        nt = L.orc_friction(len(Xb), _p(Xb), _p(X) if ev else None, _p(Xn) if ev else None, n, _p(rows), _p(w), dHat2, kappa, thickness, epsv2_h2, mu,
                            int(project_spd), C.byref(nf), _p(frows), _p(cp), _p(basis), _p(lam), C.byref(E), _p(g), cap, _p(tr), _p(tc), _p(tv))
        k = nf.value
        return dict(rows=frows[:k], closest=cp[:k], basis=basis[:k], normal_force=lam[:k], E=E.value, g=g, triplets=(tr[:nt], tc[:nt], tv[:nt]))

    def row_EgH(self, mesh, row, weight, dHat2, kappa, thickness=0.0, project_spd=True):
        m, _keep = mesh
        row = _i32(row)
        E = C.c_double(0.0); g = np.zeros(12); H = np.zeros(144); nv = C.c_int(0); verts = np.zeros(4, np.int32)
        st = self.lib.orc_row_EgH(C.byref(m), _p(row), weight, dHat2, kappa, thickness, int(project_spd), C.byref(E),
                                  _p(g), _p(H), C.byref(nv), _p(verts))
        n = 3 * nv.value
        return st, E.value, g[:n].copy(), H[:n * n].reshape(n, n).copy(), verts[:nv.value].copy()

    def min_dist2(self, mesh, rows, thickness=0.0):
        m, _keep = mesh
        rows = _i32(rows)
        d = np.empty(len(rows), np.float64); mn = C.c_double(np.nan)
        self.lib.orc_min_dist2(C.byref(m), len(rows), _p(rows), thickness, _p(d), C.byref(mn))
        return d, mn.value

    def ccd(self, mesh, direction, step=1.0, thickness=0.0, brute=False, want_cand=False):
        m, _keep = mesh
        direction = _f64(direction)
        s = C.c_double(step); sc = C.c_double(0); it = C.c_long(0); st = C.c_int(0)
        h = self.lib.orc_ccd(C.byref(m), _p(direction), thickness, int(brute), int(want_cand), C.byref(s), C.byref(sc),
                             C.byref(it), C.byref(st))
        cpt = np.empty((self.lib.orc_ccd_count(h, 1), 2), np.int32)
        cee = np.empty((self.lib.orc_ccd_count(h, 2), 2), np.int32)
        self.lib.orc_ccd_copy(h, _p(cpt), _p(cee))
        self.lib.orc_ccd_free(h)
        return {"status": st.value, "step": s.value, "step_after_clamp": sc.value, "iters": it.value,
                "cand_pt": cpt, "cand_ee": cee}

    # ---- per-pair helpers ----
    def pt_type(self, x):
        x = _f64(x); return self.lib.orc_pt_type(_p(x))

    def ee_type(self, x):
        x = _f64(x); return self.lib.orc_ee_type(_p(x))

    def dist2(self, kind, x):
        x = _f64(x); return self.lib.orc_dist2(kind, _p(x))

    def grad_hess(self, kind, x, jet=False):
        n = {0: 6, 1: 9, 2: 12, 3: 12, 6: 12}[kind]
        x = _f64(x); g = np.zeros(n); H = np.zeros((n, n))
        self.lib.orc_grad_hess(kind, int(jet), _p(x), _p(g), _p(H))
        return g, H

    def mollifier(self, x, eps_x):
        x = _f64(x); e = C.c_double(0); g = np.zeros(12); H = np.zeros((12, 12))
        self.lib.orc_mollifier(_p(x), eps_x, C.byref(e), _p(g), _p(H))
        return e.value, g, H

    def mollifier_threshold(self, x0):
        x0 = _f64(x0); return self.lib.orc_mollifier_threshold(_p(x0))

    def barrier_scalar(self, d, dHat2, kappa):
        b = C.c_double(0); g = C.c_double(0); h = C.c_double(0)
        self.lib.orc_barrier_scalar(d, dHat2, kappa, C.byref(b), C.byref(g), C.byref(h))
        return b.value, g.value, h.value

    def make_pd(self, H):
        H = _f64(H).copy(); self.lib.orc_make_pd(H.shape[0], _p(H)); return H

    def sym_eig(self, A):
        A = _f64(A); n = A.shape[0]; lam = np.zeros(n); V = np.zeros((n, n))
        self.lib.orc_sym_eig(n, _p(A), _p(lam), _p(V)); return lam, V

    def accd(self, kind, x, d, toc, eta=0.1, thickness=0.0):
        x = _f64(x); d = _f64(d); t = C.c_double(toc); it = C.c_long(0)
        hit = self.lib.orc_accd(kind, _p(x), _p(d), eta, thickness, C.byref(t), C.byref(it))
        return bool(hit), t.value, it.value

    def aabb(self, kind, x, d, dist):
        x = _f64(x); d = _f64(d if d is not None else np.zeros_like(x))
        return bool(self.lib.orc_aabb(kind, _p(x), _p(d), dist))

    def tree_mean(self, a):
        a = _f64(a); return self.lib.orc_tree_mean(_p(a), len(a))

    def mean_edge_length(self, mesh):
        m, _keep = mesh
        return self.lib.orc_mean_edge_length(C.byref(m))
