"""torchrun helper: wall-clock of each C-ABI call of one step, per rank (debugging aid for the sharded path)."""
import os, sys, time
import torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from idp_b200 import ContactContext
rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); lr = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
mesh, direction, dhat = bench.build_workload(sys.argv[1] if len(sys.argv) > 1 else "sheets8x500")
ctx = ContactContext(lr)
if world > 1:
    uid = [ctx.unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    ctx.comm_init(rank, world, uid[0])
ctx.set_surface_mesh(mesh); ctx.set_search_direction(direction)
dh2 = dhat * dhat
for it in range(4):
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
    t = [time.perf_counter()]
    ctx.constraint_set(dh2); t.append(time.perf_counter())
    ctx.barrier_all(dh2, bench.KAPPA); t.append(time.perf_counter())
    ctx.ccd_step_resident(1.0); t.append(time.perf_counter())
    ctx.min_dist2(want_all=False); t.append(time.perf_counter())
    print("rank %d it %d: ccs %.1f barrier_all %.1f ccd %.1f mind %.1f | total %.1f ms | stages %s" % (
        rank, it, *[1e3 * (t[i + 1] - t[i]) for i in range(4)], 1e3 * (t[-1] - t[0]),
        {k: round(v, 1) for k, v in ctx.stage_ms().items() if v > 0.5}), flush=True)
if world > 1:
    dist.destroy_process_group()
