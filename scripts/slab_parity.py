"""torchrun script: bench.py's N > 1 parity object alone (slab FSI step vs the single-GPU step, 1025^2, config-4 geometry)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import bench
rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
SAME_GPU = os.environ.get("RMT_SAME_GPU") == "1"      # all ranks on cuda:0, gloo for the set-up
torch.cuda.set_device(0 if SAME_GPU else local)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
if SAME_GPU:
    dist.init_process_group("gloo")
else:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1025
out = bench.slab_parity_vs_1gpu(N, "weno5", rank, world, nsteps=int(sys.argv[2]) if len(sys.argv) > 2 else 3)
if rank == 0:
    print(json.dumps(out))
dist.destroy_process_group()
