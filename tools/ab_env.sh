#!/bin/bash
# A/B of an environment switch on the SAME box: tools/ab_env.sh VAR=1 [rounds] [extra bench args]  -> alternating off / on
V=$1; R=${2:-2}; shift; shift
for i in $(seq $R); do for ON in 0 1; do
  if [ $ON = 1 ]; then export $V; TAG="$V"; else unset ${V%%=*}; TAG="default"; fi
  python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-extras "$@" 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); r=d['roofline']
print('%-22s' % '$TAG', '%.2f ms' % d['ms_per_step'], 'e2e %.2f' % d['e2e']['ms_per_step'], d['clocks']['sm_mhz'], 'MHz', {k: round(v['ms_per_step'],2) for k,v in r['kernels'].items()})"
done; done
