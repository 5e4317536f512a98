#!/bin/sh
# A/B timing of alternative builds of libpml.so on the GPU box (PML_LIBRARY selects the library):
#   tools/ab_libs.sh build/libpml_a.so build/libpml_b.so ...
for lib in "$@"; do
  echo "== $lib"
  PML_LIBRARY=$lib python bench.py --steps 200 --warmup 10 --no-cpu-baseline --no-e2e --no-extra 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('  step %.4f ms  value %.1f Mpix/s  sweep kernel %.4f ms  frac %.4f' % (d['ms_per_step'], d['value'], d['roofline']['kernel_ms'], d['roofline']['frac']))
"
done
