"""One a4 4K frame (the bench workload, 4 reference tiles) for ncu captures: python tools/prof_frame.py [frames]"""
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
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 1):
  y = runSR.sr(opt)(x)
torch.cuda.synchronize()
print('ok', tuple(y.shape))
