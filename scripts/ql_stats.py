"""Development aid: QL trip/chase statistics of the 9x9 PSD projection on real constraint rows, and the cost of a warp
running 32 consecutive rows in lock step (sum over trips of the longest chase in the warp)."""
import ctypes as C, os, subprocess, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from idp_b200 import meshgen
from oracle.binding import Oracle

out = os.path.join(ROOT, "tests", "host_shim", "libpair_host_stats.so")
subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-Wno-unknown-pragmas", "-DIDP_QL_STATS",
                       "-o", out, os.path.join(ROOT, "tests", "host_shim", "pair_host.cpp")])
hs = C.CDLL(out)
hs.hs_row_EgH.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int,
                          C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
P = lambda a: a.ctypes.data_as(C.c_void_p)
orc = Oracle()
m, d = meshgen.sheet_stack(n_sheets=4, nx=40, ny=40)
om = orc.mesh(m.X, m.X0, m.bnode, m.bedge, m.btri, m.dbc)
dh = 2e-3
rows, info, _, _ = orc.constraint_set(om, dh * dh)
four = rows[(rows[:, 0] >= 0) | (rows[:, 3] >= 0)]
print("rows", len(rows), "four-vertex", len(four))
traces = []
g = np.zeros(12); H = np.zeros(144); vv = np.zeros(4, np.int32); nv = C.c_int(0); E = C.c_double(0); buf = np.zeros(512, np.int32)
for r in four[:32 * 300]:
    r = np.ascontiguousarray(r)
    hs.hs_row_EgH(P(r), P(m.X), P(m.X0), 1.0, dh * dh, 1e5, 0.0, 1, 1, C.byref(E), P(g), P(H), C.byref(nv), P(vv))
    n = hs.hs_ql_trace(P(buf))
    traces.append(buf[:n].copy())
trips = np.array([len(t) for t in traces]); giv = np.array([t.sum() for t in traces])
print("per row: trips mean %.1f max %d; givens mean %.1f max %d" % (trips.mean(), trips.max(), giv.mean(), giv.max()))
lock = []
for w in range(0, len(traces) - 31, 32):
    T = max(len(t) for t in traces[w:w + 32])
    lock.append(sum(max((t[k] if k < len(t) else 0) for t in traces[w:w + 32]) for k in range(T)))
print("warp lock-step givens per row-slot: mean %.1f; flattened (max total) mean %.1f" % (
    np.mean(lock), np.mean([max(t.sum() for t in traces[w:w + 32]) for w in range(0, len(traces) - 31, 32)])))
