"""torchrun script: torch.profiler timeline of the slab FSI step at 4097^2 (where does the non-kernel time go?)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from torch.profiler import profile, ProfilerActivity
from pyrmt_b200.driver import make_case
from pyrmt_b200.slab import SlabFSISolver, SlabLayout
rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
torch.cuda.set_device(local)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
N = int(sys.argv[1]) if len(sys.argv) > 1 else 4097
state, prm = make_case(N, L=1.0, k_side=8, R_frac=0.04, scheme="weno5", bc_kind="lid")
lay = SlabLayout(N, N, world, rank, halo=12)
solver = SlabFSISolver(lay, prm["bc"], prm["eig"], prm["phi_init"], overlap=512, layers=prm["layers"])
state = tuple(lay.take(t).contiguous() for t in state)
sprm = dict(prm, X=None, Y=None)
def step(st):
    return solver.fsi_step(st, sprm, None, check_guard=False)
for _ in range(5):
    state = step(state)
torch.cuda.synchronize(); dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    state = step(state)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        state = step(state)
    torch.cuda.synchronize()
if rank == 0:
    print("ms_per_step", ms)
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=60))
    prof.export_chrome_trace("gpurun_out/slab_trace_r0.json")
dist.barrier()
dist.destroy_process_group()
