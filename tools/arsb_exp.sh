#!/bin/bash
# timing-only experiments on the single-CTA fused residual block (results are wrong for EXP != 0)
for e in 0 1 2 3 4 8 12 15; do
  echo "MOE_ARSB_EXP=$e"
  MOE_ARSB_EXP=$e AB_SKIP_DEFAULT=1 timeout 120 python tools/ab_flags.py arsb_solo 2>&1 | tail -1
done
