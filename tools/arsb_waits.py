"""Where the MMA warp of the single-CTA fused residual block waits (MOE_ARSB_EXP=1 switches the in-kernel clock64 accounting on):
   MOE_ARSB_EXP=1 python tools/arsb_waits.py"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import helpers as H
from moephoto_b200 import runSR, imageProcess as IP
from moephoto_b200.config import config
config.freeMemOverride = int(178 * 2 ** 30 * .9)
opt = runSR.getOpt({'model': 'a', 'scale': 4}, weights=H.load_weights('a4'))
f = runSR.sr(opt)
eng = IP.getEngine()
eng.set_conv_path(arsb_solo=True)
dbg = torch.zeros(74 * 4 + 96, dtype=torch.int64, device='cuda')
eng.debug_buffer(dbg)
x = torch.rand(3, 2160, 968, generator=torch.Generator().manual_seed(0)).half().cuda()
for _ in range(3):
  y = f(x)
torch.cuda.synchronize()
d = dbg[74 * 4:74 * 4 + 96].cpu().view(8, 2, 6)
for b in range(8):
  for w, nm in enumerate(('conv_1', 'conv_2')):
    r = d[b, w].tolist()
    n = max(1, r[0])
    print('cta %d %s steps %4d  cycles/step %6.0f | operand wait %5.0f  accumulator wait %5.0f  token wait %5.0f  MMA issue + first commit %5.0f  releases %5.0f' % (b, nm, r[0], r[1] / n, (r[2] & 0xffffffff) / n, (r[2] >> 32) / n, r[3] / n, r[4] / n, r[5] / n))
eng.debug_buffer(None)
