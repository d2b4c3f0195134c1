"""Feasibility + latency probe (2 GPUs, torchrun): one-shot reduction over peer memory (torch symmetric memory:
peers' buffers mapped over NVLink, device-side barrier on signal pads) against dist.all_reduce, for the two planes the
event-sharded objective exchanges (IWE 3.7 MB, dflow 7.4 MB).
    torchrun --nproc-per-node 2 --master-addr 127.0.0.1 profiles/tools/symm_probe.py"""
import os, time
import torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
dev = torch.device("cuda", rank)
for shape in ((720, 1280), (2, 720, 1280)):
    buf = symm.empty(shape, dtype=torch.float32, device=dev)
    hdl = symm.rendezvous(buf, dist.group.WORLD)
    peers = [hdl.get_buffer(r, shape, torch.float32) for r in range(world)]
    out = torch.empty(shape, dtype=torch.float32, device=dev)
    ref = torch.empty(shape, dtype=torch.float32, device=dev)

    def fill():
        buf.fill_(float(rank + 1))

    def one_shot():
        hdl.barrier(channel=0)                 # every rank's partial plane is complete
        torch.add(peers[0], peers[1], out=out) if world == 2 else torch.stack(peers).sum(0, out=out)
        hdl.barrier(channel=1)                 # nobody overwrites its plane while a peer still reads it

    def nccl():
        ref.copy_(buf)
        dist.all_reduce(ref)

    fill(); one_shot(); nccl(); torch.cuda.synchronize()
    assert torch.equal(out, ref), "one-shot result differs"
    for name, fn in (("one-shot peer-memory", one_shot), ("nccl all_reduce(+copy)", nccl)):
        for _ in range(5):
            fn()
        torch.cuda.synchronize(); dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(50):
            fn()
        b.record(); torch.cuda.synchronize()
        if rank == 0:
            print(f"{shape}: {name}: {a.elapsed_time(b) / 50 * 1e3:.1f} us", flush=True)
dist.destroy_process_group()
