"""A/B of the pair kernels' item scheduler on a real B200 (run under gpurun): round-robin dealing vs drawing items from
a counter.  Prints per-kernel-class milliseconds of the bench workload for both modes and, per pair of the last
pair-kernel launch, how long it ran and how many items it took.  Output -> gpurun_out/pair_balance.log"""
import os
import sys
import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import bench                                                    # noqa: E402
import tools.gpu_diag as D                                      # noqa: E402
from moephoto_b200 import runSR, imageProcess as IP            # noqa: E402
from moephoto_b200.config import config                        # noqa: E402


def say(*a):
  D.say(*a)


def spread(eng, dbg, label):
  torch.cuda.synchronize()
  d = dbg.cpu().numpy().reshape(-1, 4)
  d = d[d[:, 1] > 0]
  dur = (d[:, 1] - d[:, 0]) / 1e3
  t_end = (d[:, 1] - d[:, 0].min()) / 1e3
  say('  %s: %d pairs  run us min/avg/max %.0f/%.0f/%.0f  finish us min/max %.0f/%.0f  items min/max %d/%d' %
      (label, len(d), dur.min(), dur.mean(), dur.max(), t_end.min(), t_end.max(), d[:, 3].min(), d[:, 3].max()))
  order = np.argsort(dur)
  say('    fastest pairs (pair:sm:us:items)', ' '.join('%d:%d:%.0f:%d' % (i, d[i, 2], dur[i], d[i, 3]) for i in order[:6]))
  say('    slowest pairs (pair:sm:us:items)', ' '.join('%d:%d:%.0f:%d' % (i, d[i, 2], dur[i], d[i, 3]) for i in order[-6:]))


def main():
  config.freeMemOverride = int(178 * 2 ** 30 * .9)
  sd, _ = bench.a4_weights()
  opt = runSR.getOpt({'model': 'a', 'scale': 4}, weights=sd)
  x = IP.toTorch(8)(bench.synthetic_frame(bench.H_IN, bench.W_IN, 0))
  f = runSR.sr(opt)
  y = f(x)
  eng = opt.modelCached.engine
  ref = y.clone()
  dbg = torch.zeros(4 * 128, dtype=torch.int64, device='cuda')
  for name, static in (('round-robin', True), ('drawn', False), ('round-robin', True), ('drawn', False)):
    eng.set_conv_path(static_sched=static)
    f(x)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(3):
      y = f(x)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 3
    eng.profile(True)
    for _ in range(2):
      y = f(x)
    eng.profile(False)
    prof = eng.profile_read()
    say('%-11s %.2f ms/frame   %s   identical to first run: %s' % (
        name, ms, '  '.join('%s %.2f' % (k, prof[k][0] / 2) for k in ('conv_input', 'conv_trunk', 'conv_up', 'head')), bool(torch.equal(y, ref))))
  # per-pair run times of single launches
  g = torch.Generator().manual_seed(0)
  n, h, w = 3, 2160, 968
  xin = (torch.randn(n, h, w, 64, generator=g) * 0.5).half().cuda()
  wt = (torch.randn(64, 64, 3, 3, generator=g) * 0.05).half()
  skip = torch.randn(n, h, w, 64, generator=g).half().cuda()
  wt4 = (torch.randn(256, 64, 3, 3, generator=g) * 0.05).half()
  b4 = (torch.randn(256, generator=g) * 0.1).half()
  eng.debug_buffer(dbg)
  for name, static in (('round-robin', True), ('drawn', False)):
    eng.set_conv_path(static_sched=static)
    say(name)
    for label, args in (('trunk conv + skip (3x2160x968)', (xin, wt, None, 1, 2, 0.25, skip)),
                        ('trunk conv PReLU', (xin, wt, None, 1, 1, 0.25, None)),
                        ('upsample conv r=2 (unfused kernel)', (xin, wt4, b4, 2, 3, 0.25, None))):
      D.run_conv(eng, *args)
      dbg.zero_()
      D.run_conv(eng, *args)
      spread(eng, dbg, label)
    dbg.zero_()
    f(x)
    spread(eng, dbg, 'fused upsample+head, last launch of the frame')
  eng.debug_buffer(None)
  eng.set_conv_path()


if __name__ == '__main__':
  D.LOG = open(os.path.join(D.OUT, 'pair_balance.log'), 'w')
  main()
