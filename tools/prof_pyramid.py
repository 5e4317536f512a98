"""Run the device colour pyramid on headline-sized frames (3 frames x B=12, 192x640) for ncu / timing."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ssde_b200 import functional as Fn
N, H, W = 36, 192, 640
g = torch.Generator().manual_seed(0)
sets = [torch.randint(0, 256, (N, H, W, 3), generator=g, dtype=torch.uint8).cuda() for _ in range(8)]   # 8 x 13.3 MB
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for i in range(3):
    Fn.color_pyramid(sets[i], 4)
torch.cuda.synchronize()
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 40
e0.record()
for i in range(reps):
    Fn.color_pyramid(sets[i % len(sets)], 4)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
px = N * H * W
print("pyramid (4 launches) %.4f ms per call, %.1f GB/s algorithmic (23.8 B per scale-0 pixel)" % (ms, px * 23.8125 / ms / 1e6))
