"""Accuracy of the selected-eigenvector PSD projection prototype (scripts/proto/psd_selected.cpp) against numpy."""
import ctypes as C, os, subprocess, sys
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
so = os.path.join(HERE, "libpsd_selected.so")
subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-o", so, os.path.join(HERE, "psd_selected.cpp")])
L = C.CDLL(so)
L.psd_selected.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
P = lambda a: a.ctypes.data_as(C.c_void_p)
rng = np.random.default_rng(12)
worst = 0.0

def check(A, n):
    global worst
    A = 0.5 * (A + A.T)
    out = np.zeros((n, n)); st = C.c_long(0)
    rc = L.psd_selected(n, P(np.ascontiguousarray(A)), P(out), C.byref(st))
    lam, V = np.linalg.eigh(A)
    ref = (V * np.maximum(lam, 0.0)) @ V.T
    scale = max(np.abs(lam).max(), 1e-300)
    err = np.abs(out - ref).max() / scale
    worst = max(worst, err)
    return rc, err

for n in (6, 9):
    def spec(lam):
        Q, _ = np.linalg.qr(rng.normal(size=(n, n))); return (Q * np.asarray(lam, float)) @ Q.T
    errs = [check(rng.normal(size=(n, n)), n) for _ in range(3000)]
    print(n, "random: max err %.2e, failures %d" % (max(e for _, e in errs), sum(r for r, _ in errs)))
    hard = []
    hard.append(check(np.zeros((n, n)), n)); hard.append(check(np.eye(n), n)); hard.append(check(-np.eye(n), n))
    hard.append(check(np.diag(np.arange(n) - n / 2.0), n))
    v = rng.normal(size=n); hard.append(check(np.outer(v, v), n)); hard.append(check(-np.outer(v, v), n))
    for _ in range(200):
        lam = rng.normal(size=n); lam[1] = lam[0]; lam[3] = lam[2] * (1 + 1e-13); hard.append(check(spec(lam), n))
        hard.append(check(spec(np.concatenate([[-1.0] * (n // 2), [2.0] * (n - n // 2)])), n))
        hard.append(check(spec(np.concatenate([[1e8], rng.normal(size=n - 3) * 1e-8, [0.0, -1e-8]])), n))
        hard.append(check(spec(np.abs(rng.normal(size=n))), n))
        hard.append(check(spec(rng.normal(size=n)) * 1e-150, n)); hard.append(check(spec(rng.normal(size=n)) * 1e120, n))
        lam = -np.abs(rng.normal(size=n)); lam[:2] = np.abs(lam[:2]); hard.append(check(spec(lam), n))
    print(n, "hard spectra: max err %.2e, failures %d" % (max(e for _, e in hard), sum(r for r, _ in hard)))
print("worst", worst)
