"""torchrun --nproc-per-node N scripts/mgpu_check.py : sharded results (NCCL reductions) against the unsharded path."""
import os
import sys

import numpy as np
import scipy.sparse as sp
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from idp_b200 import ContactContext, meshgen  # noqa: E402

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
mesh, direction = meshgen.sheet_stack(n_sheets=4, nx=60, ny=50, h=4e-3, A=1.5e-3, extent=(0.12, 0.1))
dh2, kappa = (2e-3) ** 2, 1e5
ref = ContactContext(lr)
ref.set_surface_mesh(mesh)
n0 = ref.constraint_set(dh2)
rows0, _ = ref.get_constraints()
E0 = ref.barrier_energy(dh2, kappa); g0 = ref.barrier_gradient(dh2, kappa)
ptr0, col0, val0 = ref.barrier_hessian(dh2, kappa)
a0 = ref.ccd_step(direction, 1.0)
d0, m0 = ref.min_dist2()
ctx = ContactContext(lr)
uid = [ctx.unique_id() if rank == 0 else None]
dist.broadcast_object_list(uid, src=0)
ctx.comm_init(rank, world, uid[0])
ctx.set_surface_mesh(mesh)
n = ctx.constraint_set(dh2)
rows, _ = ctx.gather_constraints()
E = ctx.barrier_energy(dh2, kappa); g = ctx.barrier_gradient(dh2, kappa)
ptr, col, val = ctx.barrier_hessian(dh2, kappa)
a = ctx.ccd_step(direction, 1.0)
_, m1 = ctx.min_dist2()
_, _, d1 = ctx.gather_constraints(want_dist2=True)  # collective: per-row vector in the global order
_, m2 = ctx.min_dist2(want_all=False)  # min only: local rows + all-reduce
N = 3 * mesh.nV
parts = [None] * world
dist.all_gather_object(parts, sp.csr_matrix((val, col, ptr), shape=(N, N)))
ok = True
if rank == 0:
    H = sum(parts[1:], parts[0])
    H0 = sp.csr_matrix((val0, col0, ptr0), shape=(N, N))
    checks = {"rows": n == n0 and np.array_equal(rows, rows0), "E": abs(E - E0) <= 1e-12 * abs(E0),
              "g": np.abs(g - g0).max() <= 1e-12 * np.abs(g0).max(), "H": abs(H - H0).max() <= 1e-12 * abs(H0).max(),
              "alpha": a == a0, "dist2": np.array_equal(d1, d0) and m1 == m0 and m2 == m0}
    print("mgpu_check world=%d rows=%d/%d E=%.15e/%.15e alpha=%.15e/%.15e %s" % (world, n, n0, E, E0, a, a0, checks), flush=True)
    ok = all(checks.values())
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
