"""Timings of the other BASELINE.json configs on one B200 (parity-test cases, not bench lines): python tools/bench_configs.py"""
import os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import helpers as H
import bench
from moephoto_b200 import runSR, runDN, imageProcess as IP, video
from moephoto_b200.config import config

config.freeMemOverride = int(178 * 2 ** 30 * .9)


def timed(fn, reps=3):
  fn(); torch.cuda.synchronize()
  a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  a.record()
  for _ in range(reps):
    r = fn()
  b.record(); torch.cuda.synchronize()
  return a.elapsed_time(b) / reps, r


# configs[0]: single 256x256 tile, a2
o2 = runSR.getOpt({'model': 'a', 'scale': 2}, weights=H.load_weights('a2'))
x = IP.toTorch(8)(bench.synthetic_frame(256, 256, 1))
ms, y = timed(lambda: runSR.sr(o2)(x), 10)
print('configs[0] a2 256x256 -> 512x512: %.3f ms  %.1f MPix/s out' % (ms, 0.262144 / ms * 1e3))
# configs[1]: 1080p -> 4K a2
x = IP.toTorch(8)(bench.synthetic_frame(1080, 1920, 2))
ms, y = timed(lambda: runSR.sr(o2)(x), 5)
print('configs[1] a2 1920x1080 -> 3840x2160 (%d tile): %.2f ms  %.1f MPix/s out' % (len(o2.plan.tiles), ms, 8.2944 / ms * 1e3))
# configs[3]: dn_lite15 -> a2 chained, 1080p, batch of 16 frames (host uint8 frames in, host frames out)
odn = runDN.getOpt({'model': 'lite15'}, weights=H.load_weights('dn_lite15'))
frames = [bench.synthetic_frame(1080, 1920, 10 + i) for i in range(16)]
fb = video.FrameBatcher([odn, o2], 1080, 1920, bit_depth=8, swap_rb=False, batch=8)
out = dict(fb.process(frames, copy=False))          # warm-up (workspace, plans)
torch.cuda.synchronize(); t1 = time.perf_counter()
n_out = sum(1 for _ in fb.process(frames, copy=False))
torch.cuda.synchronize(); t2 = time.perf_counter()
print('configs[3] dn_lite15 -> a2, 16 x 1080p frames (8 per engine call), host in/out: %.1f ms total = %.2f ms/frame  %.1f MPix/s out'
      % ((t2 - t1) * 1e3, (t2 - t1) * 1e3 / 16, 16 * 8.2944 / (t2 - t1)))
del out
# configs[4] in miniature: a3 on 4K frames (the full config is 64 frames over 8 GPUs: frames are independent)
o3 = runSR.getOpt({'model': 'a', 'scale': 3}, weights=H.load_weights('a3'))
x = IP.toTorch(8)(bench.synthetic_frame(2160, 3840, 3))
ms, y = timed(lambda: runSR.sr(o3)(x), 3)
print('configs[4] a3 3840x2160 -> 11520x6480 (%d tiles): %.2f ms/frame  %.1f MPix/s out  (64 frames on 8 GPUs = 8 frames each)' % (len(o3.plan.tiles), ms, 74.6496 / ms * 1e3))
