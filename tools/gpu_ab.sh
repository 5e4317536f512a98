#!/bin/sh
# A/B timing of the fused sweep on the GPU box: env-var variants of one built library.
#   tools/gpu_ab.sh "PML_PREFETCH=0" "PML_PREFETCH=1" ...
for v in "$@"; do
  echo "== $v"
  env $v python bench.py --steps 200 --warmup 10 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('  step %.4f ms  value %.1f Mpix/s  kernel %.4f ms  frac %.4f' % (d['ms_per_step'], d['value'], d['roofline']['kernel_ms'], d['roofline']['frac']))
"
done
