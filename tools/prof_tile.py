"""One reference tile of the bench workload (a4 on 3 x 2160 x 968, what one of the four column strips of a 4K frame is) for
ncu captures:  [MOE_CONV_FLAGS=arsb_solo,...] python tools/prof_tile.py [model key] [scale] [h] [w] [runs]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import helpers as H
from moephoto_b200 import runSR, runDN, imageProcess as IP
from moephoto_b200.config import config
key = sys.argv[1] if len(sys.argv) > 1 else 'a4'
scale = int(sys.argv[2]) if len(sys.argv) > 2 else 4
h = int(sys.argv[3]) if len(sys.argv) > 3 else 2160
w = int(sys.argv[4]) if len(sys.argv) > 4 else 968
runs = int(sys.argv[5]) if len(sys.argv) > 5 else 2
config.freeMemOverride = int(178 * 2 ** 30 * .9)
sd = H.load_weights(key)
if key.startswith('dn'):
  opt = runDN.getOpt({'model': key.split('_')[1]}, weights=sd)
  f = IP.RGBFilter(opt)
else:
  opt = runSR.getOpt({'model': 'lite' if key.startswith('lite') else key[0], 'scale': scale}, weights=sd)
  f = runSR.sr(opt)
IP.getEngine().set_conv_path(**{k: True for k in os.environ.get('MOE_CONV_FLAGS', '').split(',') if k})
x = torch.rand(3, h, w, generator=torch.Generator().manual_seed(0)).half().cuda()
for _ in range(runs):
  y = f(x)
torch.cuda.synchronize()
print('ok', key, tuple(y.shape), len(opt.plan.tiles), 'tiles', IP.getEngine().launches(), 'launches')
