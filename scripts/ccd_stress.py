"""BASELINE configs[4]: CCD-only stress on the 16 M-triangle near-contact sheets, step-size
filter sweep. Prints one JSON line per (sigma, alpha0, thickness) with candidate pairs/s of the resident CCD path
(idp_ccd_step_resident: broad phase on swept boxes + additive CCD + min over candidates), device-timed.

    python scripts/ccd_stress.py                       # 1 GPU
    python -m torch.distributed.run --nproc-per-node N scripts/ccd_stress.py   # N GPUs (sharded queries, all-reduce(min))
"""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from idp_b200 import ContactContext, meshgen  # noqa: E402

rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); lr = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
h = 2e-3
t0 = time.time()
mesh, d = meshgen.sheet_stack(n_sheets=16, nx=1000, ny=500, h=h, A=0.75e-3, seed=20260104, dir_sigma=1.0, dir_seed=20260105, extent=(2.0, 1.0))
d[:, 2] -= np.where((np.arange(len(d)) // (1001 * 501)) % 2 == 1, -1.0, 1.0) * 0.25 * h  # unit Gaussian direction
gen_s = time.time() - t0
ctx = ContactContext(lr)
if world > 1:
    uid = [ctx.unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    ctx.comm_init(rank, world, uid[0])
ctx.set_surface_mesh(mesh)
for sigma in (0.5 * h, 1.0 * h, 2.0 * h, 4.0 * h):
    ctx.set_search_direction(np.ascontiguousarray(d * sigma))
    for a0 in (1.0, 0.5, 0.25, 0.1):  # the full alpha0 grid
        for xi in (0.0, 1e-4):
            for _ in range(2):
                a = ctx.ccd_step_resident(a0, xi)  # warm-up (buffers sized)
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 3
            e0.record()
            for _ in range(reps):
                a = ctx.ccd_step_resident(a0, xi)
            e1.record()
            torch.cuda.synchronize()
            ms = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device="cuda")
            cand = torch.tensor([float(ctx.count(3) + ctx.count(4))], dtype=torch.float64, device="cuda")
            if world > 1:
                dist.all_reduce(ms, op=dist.ReduceOp.MAX)
                dist.all_reduce(cand, op=dist.ReduceOp.SUM)
            if rank == 0:
                print(json.dumps({"workload": "sheets16x1000x500 (16,000,000 triangles), CCD only", "n_gpus": world, "sigma_over_h": sigma / h,
                                  "alpha0": a0, "thickness": xi, "alpha": a, "ccd_candidates": int(cand.item()), "ms": ms.item(),
                                  "pairs_per_s": cand.item() / (ms.item() * 1e-3), "mesh_gen_s": round(gen_s, 1)}), flush=True)
ctx.close()
if world > 1:
    dist.destroy_process_group()
