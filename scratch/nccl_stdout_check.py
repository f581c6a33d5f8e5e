"""GPU check (not product code): the stdout redirect bench.py uses around NCCL start-up, under torchrun."""
import os, sys, json
import torch, torch.distributed as dist
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); device = torch.device("cuda", lr)
sys.stdout.flush()
try:
    saved = os.dup(1); os.dup2(2, 1)
except OSError:
    saved = None
try:
    dist.init_process_group("nccl", device_id=device); dist.barrier(); torch.cuda.synchronize(device)
finally:
    if saved is not None:
        sys.stdout.flush(); os.dup2(saved, 1); os.close(saved)
t = torch.ones(4, device=device); dist.all_reduce(t)
if rank == 0:
    print(json.dumps({"ok": True, "sum": float(t[0])}), flush=True)
dist.barrier(); dist.destroy_process_group()
