"""Probe: does torch symmetric memory (peer-mapped buffers over NVLink) work on this box?
usage: python -m torch.distributed.run --nproc-per-node N scripts/symm_probe.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem
from prospect_public_b200.distributed import init_from_env
rank, world, _ = init_from_env('nccl')
dev = torch.device(f'cuda:{torch.cuda.current_device()}')
t = symm_mem.empty(1024, dtype=torch.float64, device=dev)
t.fill_(-1.0)
hdl = symm_mem.rendezvous(t, dist.group.WORLD)
print(rank, 'rendezvous ok; attrs', [a for a in dir(hdl) if not a.startswith('_')][:40], flush=True)
ptrs = list(hdl.buffer_ptrs)
print(rank, 'buffer_ptrs', [hex(p) for p in ptrs], 'local', hex(t.data_ptr()), flush=True)
hdl.barrier()
for p in range(world):                       # write my rank id into slot `rank` of every peer's buffer
    peer = hdl.get_buffer(p, (1024,), torch.float64)
    peer[rank] = float(rank) + 0.5
torch.cuda.synchronize()
hdl.barrier()
torch.cuda.synchronize()
print(rank, 'after exchange', t[:world].tolist(), flush=True)
assert t[:world].tolist() == [r + 0.5 for r in range(world)]
dist.barrier()
dist.destroy_process_group()
