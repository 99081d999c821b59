"""Global batch 256 split over the ranks (strong scaling), eager DDP next to the CUDA-graph step with the gradient all-reduce
captured inside -- bench.py's `secondary.mae_global256` on its own:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29512 tools/global256.py"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import bench
import headct_foundation_b200 as H
from headct_foundation_b200.optim import FusedAdamW

world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
out = bench.secondary_global256(torch, dist, H, FusedAdamW, dev, world, local, bench._peaks()["tflops"])
if rank == 0:
    print(json.dumps({"n_gpus": world, **out}))
if world > 1:
    dist.destroy_process_group()
