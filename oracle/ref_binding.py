"""ORACLE — TEST INFRASTRUCTURE ONLY. ctypes view of oracle/_ref/libidp_ref.so: the reference's own per-pair math
(Library/Math/Distance/*.h, BARRIER.h, UTILS.h) compiled from /root/reference by oracle/ref_shim/Makefile."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(_HERE, "_ref", "libidp_ref.so")


def available():
    return os.path.exists(LIB)


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
