"""Which ATen op issues the stray per-step copy kernel?  (torch profiler over three fused steps)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
class A: pass
a = A(); a.batch, a.height, a.width, a.sources = 12, 192, 640, 2
dev = torch.device("cuda", 0)
opt, srcs, sets = bench.make_sets(a, 1, 0)
step = bench.fused_step_fn(opt, srcs, sets[0][0], sets[0][1], dev)
for _ in range(3): step()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], with_stack=True) as prof:
    step(); torch.cuda.synchronize()
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA: continue
    if "copy" in e.name or "clone" in e.name or "contiguous" in e.name or "to" == e.name[-2:]:
        print(e.name, e.input_shapes if hasattr(e, "input_shapes") else "", [s for s in (e.stack or [])][:4])
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=14))
