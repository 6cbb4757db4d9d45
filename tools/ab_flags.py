"""A/B of engine switches inside ONE process on one B200 (same box, same clocks): a4 on the 4K bench frame, alternating.
   python tools/ab_flags.py arsb_smem_mid [no_arsb ...]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import bench
from moephoto_b200 import runSR, imageProcess as IP
from moephoto_b200.config import config
config.freeMemOverride = int(178 * 2 ** 30 * .9)
sd, _ = bench.a4_weights()
opt = runSR.getOpt({'model': 'a', 'scale': 4}, weights=sd)
x = IP.toTorch(8)(bench.synthetic_frame(bench.H_IN, bench.W_IN, 0))
eng = IP.getEngine(0)
f = runSR.sr(opt)
ref = None
for rnd in range(2):
  for flag in ([] if os.environ.get('AB_SKIP_DEFAULT') else [None]) + sys.argv[1:]:
    eng.set_conv_path(**({flag: True} if flag else {}))
    for _ in range(2):
      y = f(x)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
      y = f(x)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 5
    eng.profile(True)
    for _ in range(2):
      y = f(x)
    eng.profile(False)
    pr = eng.profile_read()
    if ref is None:
      ref = y.clone()
    print('%-16s %.2f ms/frame  %s  identical to the default: %s' % (flag or 'default', ms, '  '.join('%s %.2f' % (k, pr[k][0] / 2) for k in ('arsb', 'conv_trunk', 'conv_up', 'conv_up_head', 'head', 'conv_input')),
                                                                   bool(torch.equal(y, ref))), flush=True)
eng.set_conv_path()
