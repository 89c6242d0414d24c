"""Runs `--steps` synthetic Newton iterations of the hot path (after `--warmup`) for ncu / timing experiments."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from idp_b200 import ContactContext  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("workload", nargs="?", default="sheets8x160")
ap.add_argument("--steps", type=int, default=1)
ap.add_argument("--warmup", type=int, default=1)
ap.add_argument("--only", default="all", choices=["all", "barrier", "ccs", "ccd"])
args = ap.parse_args()
mesh, direction, dhat = bench.build_workload(args.workload)
ctx = ContactContext(0)
ctx.set_surface_mesh(mesh)
ctx.set_search_direction(direction)
dhat2 = dhat * dhat
if args.only == "barrier":
    ctx.constraint_set(dhat2)
for it in range(args.warmup + args.steps):
    print('step', it, 'allocs so far', ctx.count(8))
    if args.only in ("all", "ccs"):
        n = ctx.constraint_set(dhat2)
    if args.only in ("all", "barrier"):
        E, nnz = ctx.barrier_all(dhat2, bench.KAPPA)
    if args.only in ("all", "ccd"):
        a = ctx.ccd_step_resident(1.0)
    if args.only == "all":
        ctx.min_dist2(want_all=False)
print({k: round(v, 3) for k, v in ctx.stage_ms().items()}, ctx.count(0), ctx.launches(), 'allocs', ctx.count(8))
