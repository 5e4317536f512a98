#!/bin/sh
# A/B timing of environment-knob variants of the built library on the GPU box:  tools/ab_env.sh "PML_PAIR=0" "PML_PAIR=1" ...
for v in "$@"; do
  echo "== $v"
  env $v python bench.py --steps 200 --warmup 10 --no-cpu-baseline --no-e2e --no-extra 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('  step %.4f ms  value %.1f Mpix/s  sweep kernel %.4f ms  frac %.4f' % (d['ms_per_step'], d['value'], d['roofline']['kernel_ms'], d['roofline']['frac']))
"
done
