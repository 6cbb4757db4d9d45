#!/bin/bash
# A/B of two builds of the library on the SAME box (box-to-box variation of the power-capped clocks is +-5 %):
#   tools/ab_bench.sh libA.so libB.so [rounds]     -> ms per frame, clocks, per-kernel-class times, alternating A B A B ...
A=$1; B=$2; R=${3:-2}
for i in $(seq $R); do for L in $A $B; do
  MOE_B200_LIB=$PWD/$L python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-extras 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); r=d['roofline']
print('%-50s' % '$L', '%.2f ms' % d['ms_per_step'], 'e2e %.2f' % d['e2e']['ms_per_step'], d['clocks']['sm_mhz'], 'MHz', {k: round(v['ms_per_step'],2) for k,v in r['kernels'].items()})"
done; done
