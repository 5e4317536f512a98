"""Run a few eager fused steps of the headline workload (for ncu / compute-sanitizer)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

class A: pass
a = A(); a.batch, a.height, a.width, a.sources = int(os.environ.get("B", 12)), int(os.environ.get("H", 192)), int(os.environ.get("W", 640)), int(os.environ.get("S", 2))
dev = torch.device("cuda", 0)
opt, srcs, sets = bench.make_sets(a, 1, 0)
step = bench.fused_step_fn(opt, srcs, sets[0][0], sets[0][1], dev)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 5
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for i in range(n):
    if i == n - 1: e0.record()
    step()
e1.record()
torch.cuda.synchronize()
print("last step ms", e0.elapsed_time(e1))
