"""Does torch symmetric memory hand out an NVLS multicast mapping here? (diagnostic)"""
import os
import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as sm

os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
os.environ.setdefault("MASTER_PORT", "29533")
rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", torch.cuda.current_device()))
t = sm.empty(1 << 20, dtype=torch.float32, device="cuda")
h = sm.rendezvous(t, dist.group.WORLD)
print("rank", rank, "world", world, "multicast_ptr", hex(int(h.multicast_ptr or 0)), flush=True)
dist.destroy_process_group()
