"""torchrun script: the periodic fluid step on P GPUs against the single-GPU operators at N x N (default 16385)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
SAME_GPU = os.environ.get("RMT_SAME_GPU") == "1"
torch.cuda.set_device(0 if SAME_GPU else local)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
dist.init_process_group("gloo") if SAME_GPU else dist.init_process_group("nccl", device_id=torch.device("cuda", local))
from pyrmt_b200.slab import periodic_fluid_parity
N = int(sys.argv[1]) if len(sys.argv) > 1 else 16385
t0 = time.time()
out = periodic_fluid_parity(N, world, rank, steps=2)
out["wall_s"] = time.time() - t0
if rank == 0:
    print(json.dumps(out))
dist.destroy_process_group()
