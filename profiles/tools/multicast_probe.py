"""Does torch symmetric memory expose a multicast (NVLS) mapping on this box?  torchrun --nproc-per-node N this file."""
import os
import torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
t = symm.empty((1 << 20,), dtype=torch.float32, device="cuda")
h = symm.rendezvous(t, dist.group.WORLD)
mc = getattr(h, "multicast_ptr", None)
print(f"rank {rank}/{world}: multicast_ptr = {mc}, has_multicast_support = "
      f"{getattr(symm, 'has_multicast_support', lambda *a: 'n/a')('cuda', torch.cuda.current_device()) if hasattr(symm, 'has_multicast_support') else 'n/a'}", flush=True)
if mc:
    t.fill_(rank + 1.0)
    h.barrier()
    try:
        out = torch.ops.symm_mem.multimem_all_reduce_(t, "sum", dist.group.WORLD.group_name)
        torch.cuda.synchronize()
        print(f"rank {rank}: multimem_all_reduce_ -> {float(t[0])} (expected {world * (world + 1) / 2})", flush=True)
    except Exception as e:
        print(f"rank {rank}: multimem_all_reduce_ failed: {e}", flush=True)
dist.destroy_process_group()
