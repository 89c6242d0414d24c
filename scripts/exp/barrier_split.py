"""Times the k_barrier launches of the bench workload in its variants (E only, E+g, H without / with the PSD projection)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench
from idp_b200 import ContactContext
wl = sys.argv[1] if len(sys.argv) > 1 else "sheets8x500"
mesh, direction, dhat = bench.build_workload(wl)
ctx = ContactContext(0)
ctx.set_surface_mesh(mesh)
d2 = dhat * dhat
ctx.constraint_set(d2)
def t(name, f, n=3):
    best = 1e9
    for _ in range(n):
        f()
        best = min(best, ctx.stage_ms()["k_barrier"])
    print("%-28s k_barrier %.3f ms" % (name, best), flush=True)
t("E", lambda: ctx.barrier_energy(d2, bench.KAPPA))
t("g", lambda: ctx.L.idp_barrier_gradient(ctx.h, d2, bench.KAPPA, 0.0, None, 3))
t("H no projection", lambda: ctx.barrier_hessian(d2, bench.KAPPA, project_spd=False, fetch=False))
t("H + PSD", lambda: ctx.barrier_hessian(d2, bench.KAPPA, project_spd=True, fetch=False))
t("E+g+H+PSD", lambda: ctx.barrier_all(d2, bench.KAPPA))
