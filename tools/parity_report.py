"""CPU only: the oracle's three fp16 modes against the committed goldens (the UNMODIFIED reference in its GPU fp16
configuration, executed on CPU).  python tools/parity_report.py > profiles/r02_parity_oracle_vs_reference.txt"""
import os
import sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'tests')]
import helpers as H  # noqa: E402

print('# oracle/net.py (numpy + C restatement) vs tests/golden/cases.npz <case>.ref16 = the UNMODIFIED reference in its GPU fp16')
print('# configuration, executed on CPU.  ref16cpu = per-op rounding, bias inside the convolution (what the CPU executes; the pinned mode);')
print('# ref16 = the same with the bias add as a separate op (what aten + cuDNN execute on the GPU; the engine contract);')
print('# f16io = round-1 contract (one rounding per stored tensor).')
for name in H.case_names():
  c = H.load_case(name)
  row = '%-16s' % name
  for mode in ('ref16cpu', 'ref16', 'f16io'):
    y = H.run_case_oracle(c, mode=mode)
    d = np.abs(y - c['ref16'])
    row += ' | %s: max %.2e mean %.2e >1e-3 %.1e PSNR %.1f dB' % (mode, d.max(), d.mean(), (d > 1e-3).mean(), H.psnr(y, c['ref16']))
  print(row + ' | golden fp16 vs golden fp32: PSNR %.1f dB' % H.psnr(c['ref16'], c['ref']), flush=True)
