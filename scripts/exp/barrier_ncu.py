import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench
from idp_b200 import ContactContext
mesh, direction, dhat = bench.build_workload("sheets8x500")
ctx = ContactContext(0)
ctx.set_surface_mesh(mesh)
d2 = dhat * dhat
ctx.constraint_set(d2)
ctx.barrier_hessian(d2, bench.KAPPA, project_spd=False, fetch=False)
ctx.barrier_hessian(d2, bench.KAPPA, project_spd=True, fetch=False)
