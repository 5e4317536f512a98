"""BASELINE.json configs[4]: loss-kernel sweep over batch, resolution and number of source frames against the
HBM roofline.  Device-resident fwd+bwd (eager launches, CUDA events around whole steps, median of `reps`),
algorithmic bytes per SURVEY.md section 8(d).  Writes one table line per configuration.
    python tools/sweep_c5.py [reps] > profiles/r01_c5_sweep.txt"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from ssde_b200 import synthetic

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 12
dev = torch.device("cuda", 0)
peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"] \
    if os.path.exists("MEASURED_PEAKS.json") else 6548.8
print("# B | HxW | S | ms/step | Mpix/s | algorithmic GB/s | fraction of %.1f GB/s (measured copy peak) | input set MB" % peak)
class A: pass
cases = []
for (H, W) in ((96, 320), (192, 640), (320, 1024), (384, 1280)):
    for S in (2, 4, 8):
        for B in (1, 4, 12, 32, 64, 128, 256):
            pix = B * H * W
            if pix * (3 * (1 + S) + 2) * 4 > 12e9 or pix > 64e6:     # keep host generation + device memory bounded
                continue
            if S > 2 and B not in (1, 12, 64):
                continue
            cases.append((B, H, W, S))
for (B, H, W, S) in cases:
    a = A(); a.batch, a.height, a.width, a.sources = B, H, W, S
    n_sets = 2 if B * H * W * (1 + S) * 12 < 400e6 else 1
    try:
        opt, srcs, sets = bench.make_sets(a, n_sets, 0)
        steps = [bench.fused_step_fn(opt, srcs, i, o, dev) for (i, o) in sets]
        for st in steps:
            st()
        torch.cuda.synchronize()
        ts = []
        for r in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); steps[r % len(steps)](); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        ms = ts[len(ts) // 2]
        ab = synthetic.algorithmic_bytes(B, H, W, S, 4)
        set_mb = B * H * W * 4 * (3 * (1 + S) + 1.33 + 3 * 0.33) / 1e6
        print("%4d | %4dx%-4d | %d | %8.3f | %8.1f | %8.1f | %.3f | %.0f" % (B, H, W, S, ms, B * H * W / ms / 1e3, ab / ms / 1e6, ab / ms / 1e6 / peak, set_mb), flush=True)
        del steps, sets
        torch.cuda.empty_cache()
    except Exception as ex:   # noqa
        print("%4d | %4dx%-4d | %d | failed: %s" % (B, H, W, S, str(ex)[:80]), flush=True)
