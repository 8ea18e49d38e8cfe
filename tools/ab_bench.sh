#!/bin/bash
# A/B on ONE box: alternate two environment settings, several rounds each (clocks drift with the power cap).
# usage: tools/ab_bench.sh "CAZ_L2_PREFETCH=0" "CAZ_L2_PREFETCH=1" [rounds]
A="$1"; B="$2"; R="${3:-3}"
for i in $(seq 1 $R); do
  for cfg in "$A" "$B"; do
    env $cfg python bench.py --no-cpu-baseline --no-latency --steps 80 --warmup 10 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read())
s = d['roofline']['stage_ms_per_step']
print('%-28s value %9.0f  e2e %9.0f  stem %.3f tower %.3f heads %.3f  clk %s' % ('$cfg', d['value'], d['e2e']['value'], s['stem'], s['tower'], s['heads'], d['clocks']['sm_mhz']))"
  done
done
