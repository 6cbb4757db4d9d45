"""Per-model timings and per-kernel-class breakdown on one B200: python tools/time_models.py  (-> profiles/r02_all_models.log)"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import helpers as H
from moephoto_b200 import runSR, runDN, imageProcess as IP
from moephoto_b200.config import config

config.freeMemOverride = int(178 * 2 ** 30 * .9)
eng = IP.getEngine(0)
# algorithmic FLOP per LR pixel-plane; lite: 1x1 convs counted with one tap
FLOP = {'a2': 1558656, 'a3': 2307456, 'a4': 3945600, 'dn_lite15': 541728}
for key, scale, shape in (('a2', 2, (3, 1080, 1920)), ('a3', 3, (3, 2160, 3840)), ('a4', 4, (3, 2160, 3840)), ('dn_lite15', 1, (3, 1080, 1920)),
                          ('dn_lite15', 1, (48, 1080, 1920)), ('lite2', 2, (3, 1080, 1920)), ('lite4', 4, (3, 1080, 1920)), ('lite8', 8, (3, 1080, 1920))):
  sd = H.load_weights(key)
  if key.startswith('dn'):
    opt = runDN.getOpt({'model': 'lite15'}, weights=sd)
    run = IP.RGBFilter(opt)
  else:
    opt = runSR.getOpt({'model': 'lite' if key.startswith('lite') else 'a', 'scale': scale}, weights=sd)
    run = runSR.sr(opt)
  x = torch.rand(shape, device='cuda').half()
  for _ in range(2):
    y = run(x)
  torch.cuda.synchronize()
  t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  t0.record()
  for _ in range(3):
    y = run(x)
  t1.record(); torch.cuda.synchronize()
  ms = t0.elapsed_time(t1) / 3
  eng.profile(True)
  for _ in range(2):
    y = run(x)
  eng.profile(False)
  pr = eng.profile_read()
  line = '%-10s %-16s tiles=%-2d %8.2f ms  %7.1f MPix/s out' % (key, shape, len(opt.plan.tiles), ms, shape[0] / 3 * y.shape[1] * y.shape[2] / ms / 1e3)
  if key in FLOP:
    line += '  %6.0f TFLOP/s algorithmic' % (FLOP[key] * shape[0] * shape[1] * shape[2] / (ms * 1e-3) / 1e12)
  print(line)
  print('           ' + '  '.join('%s %.2f ms x%d' % (k, v[0] / 2, v[2] // 2) for k, v in pr.items() if v[2] and k in IP.Engine.PROFILE_CLASSES), flush=True)
  del y, x, opt
  torch.cuda.empty_cache()
