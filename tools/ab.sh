#!/bin/bash
# A/B timing of tuning builds on the benchmark configuration (run under gpurun):
#   bash tools/ab.sh tune/a.so tune/b.so ...   -> one line per library: value, ms/step, roofline frac
for lib in "$@"; do
  BLMX_LIB=$PWD/$lib timeout 300 python bench.py --steps 2 --warmup 2 --profile ${BENCH_ARGS:-} 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); r = d['roofline']
print('$lib', '%.4g' % d['value'], '%.1f ms' % d['ms_per_step'], 'frac %.3f' % r['frac'], 'kernel %.2f ms' % r['kernel_ms_per_launch'])"
done
