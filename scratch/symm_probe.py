"""2-GPU probe: does torch's symmetric memory give peer-mapped pointers here, and do plain kernels see stores through them?"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
torch.cuda.set_device(rank)
dev = torch.device('cuda', rank)
dist.init_process_group('nccl', device_id=dev)
import torch.distributed._symmetric_memory as symm_mem
try:
    t = symm_mem.empty(1 << 20, dtype=torch.float32, device=dev)
    hdl = symm_mem.rendezvous(t, group=dist.group.WORLD)
    print(rank, 'rendezvous ok', type(hdl).__name__, [a for a in dir(hdl) if not a.startswith('_')][:40], flush=True)
    print(rank, 'buffer_ptrs', [hex(p) for p in hdl.buffer_ptrs], 'signal', [hex(p) for p in hdl.signal_pad_ptrs], 'multicast', hex(getattr(hdl, 'multicast_ptr', 0) or 0), flush=True)
    t.fill_(float(rank + 1))
    dist.barrier(); torch.cuda.synchronize()
    peer = hdl.get_buffer((rank + 1) % world, (16,), torch.float32)
    print(rank, 'peer view reads', peer[:4].tolist(), flush=True)
    # write into the peer through the mapped pointer with our own kernel (moments kernel: mean of S=1 rows = copy)
    import arpde_b200._lib as L
    src = torch.full((1, 256), float(10 * (rank + 1)), device=dev)
    peer_ptr = hdl.buffer_ptrs[(rank + 1) % world] + 4096
    L.call('arpde_predictive_moments', L.ptr(src), 256, 256, 1, 1, 0.0, peer_ptr, 0, L.stream())
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    print(rank, 'own buffer after the peer wrote', t[1024:1028].tolist(), flush=True)
except Exception as e:
    import traceback; traceback.print_exc()
    print(rank, 'FAILED', repr(e)[:300], flush=True)
dist.destroy_process_group()
