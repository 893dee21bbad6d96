"""Per-exchange latency of the peer-mailbox stage alone (run under torchrun with >= 2 GPUs)."""
import ctypes, os, sys
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gradient_free_mcmc_postprocessing_b200 import _lib, distributed as gdist

torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', 0)))
dist.init_process_group('nccl')
gdist.init(p2p=True)
lib = _lib.load()
for iters in (1000, 20000, 20000):
    us = ctypes.c_double()
    dist.barrier()
    _lib.check(lib.gfk_p2p_selftest(iters, ctypes.byref(us), None))
    if dist.get_rank() == 0:
        print(f'p2p mailbox exchange: world {dist.get_world_size()} iters {iters}: {us.value:.3f} us per exchange', flush=True)
# NCCL 16-byte all-gather for comparison (what the fallback path pays per step, plus its launch)
t = torch.zeros(2, dtype=torch.float64, device='cuda')
out = torch.zeros(2 * dist.get_world_size(), dtype=torch.float64, device='cuda')
for _ in range(20):
    dist.all_gather_into_tensor(out, t)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(2000):
    dist.all_gather_into_tensor(out, t)
e1.record(); torch.cuda.synchronize()
if dist.get_rank() == 0:
    print(f'nccl all_gather 16 B/rank: {e0.elapsed_time(e1) / 2000 * 1e3:.3f} us per call (back-to-back launches)', flush=True)
gdist.finalize(); dist.destroy_process_group()
