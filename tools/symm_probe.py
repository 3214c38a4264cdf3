"""Scratch: does torch symmetric memory work on this box? (torchrun --nproc-per-node 2 tools/symm_probe.py)"""
import os, time, torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm
rank = int(os.environ['RANK']); world = int(os.environ['WORLD_SIZE'])
torch.cuda.set_device(rank)
dist.init_process_group('nccl', device_id=torch.device('cuda', rank))
t = symm.empty((world * 4, 24), dtype=torch.float32, device=torch.device('cuda', rank))
t.zero_()
hdl = symm.rendezvous(t, dist.group.WORLD.group_name)
print(rank, 'rendezvous ok', type(hdl).__name__, [hex(p) for p in hdl.buffer_ptrs], 'signal pads', len(hdl.signal_pad_ptrs), flush=True)
# each rank writes its rows into every peer's buffer through get_buffer
for r in range(world):
    peer = hdl.get_buffer(r, (world * 4, 24), torch.float32)
    peer[rank * 4:(rank + 1) * 4] = float(rank + 1)
hdl.barrier()
torch.cuda.synchronize()
print(rank, 'rows', t[:, 0].tolist(), flush=True)
# latency of barrier vs all_gather
x = torch.ones((8192, 24), device='cuda'); out = torch.empty((world * 8192, 24), device='cuda')
for name, fn in (('symm barrier', lambda: hdl.barrier()), ('nccl all_gather 786KB', lambda: dist.all_gather_into_tensor(out, x))):
    for _ in range(5): fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50): fn()
    e1.record(); torch.cuda.synchronize()
    if rank == 0: print(name, e0.elapsed_time(e1) / 50 * 1e3, 'us', flush=True)
dist.destroy_process_group()
