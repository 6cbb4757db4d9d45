"""Diagnostics on a real B200 (run under gpurun).  Prints, never asserts: how single convolution layers compare with
the on-device SIMT cross-check and with a torch fp32 reference, golden parity on both conv paths, and per-kernel-class
timings of every model family.  Output -> gpurun_out/diag.log (copies of interesting runs live in profiles/).
"""
import ctypes
import json
import os
import sys
import time
import traceback

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from moephoto_b200 import _lib, weights as W, imageProcess as IP, runSR, runDN   # noqa: E402
from moephoto_b200.config import config   # noqa: E402

OUT = os.path.join(ROOT, 'gpurun_out')
os.makedirs(OUT, exist_ok=True)
LOG = open(os.path.join(OUT, 'diag.log'), 'w')


def say(*a):
  s = ' '.join(str(x) for x in a)
  print(s, flush=True)
  LOG.write(s + '\n')
  LOG.flush()


def conv_ref(x, w, bias, r, epi, param, skip):
  """torch fp32 reference of moe_conv3x3_c64 on NHWC fp16 data"""
  xn = x.float().permute(0, 3, 1, 2)
  y = torch.nn.functional.conv2d(xn, w.float(), None if bias is None else bias.float(), padding=1)
  if epi == 1:
    y = torch.where(y >= 0, y, param * y)
  elif epi == 2:
    y = skip.float().permute(0, 3, 1, 2) + param * y
  elif epi == 3:
    y = torch.nn.functional.pixel_shuffle(y, r) if r > 1 else y
    y = torch.where(y >= 0, y, param * y)
  return y.permute(0, 2, 3, 1).contiguous()


def run_conv(eng, x, w, bias, r, epi, param, skip):
  n, h, wd, _ = x.shape
  w16 = w.half().cpu().numpy()
  imgs, bs = [], []
  for i in range(r):
    for j in range(r):
      sel = np.arange(64) * r * r + i * r + j
      imgs.append(W.conv_image(w16[sel]))
      if bias is not None:
        bs.append(bias.float().cpu().numpy()[sel])
  img = torch.from_numpy(np.concatenate(imgs)).cuda()
  bdev = torch.from_numpy(np.stack(bs).astype(np.float32)).cuda() if bias is not None else None
  out = torch.empty((n, h * r, wd * r, 64), dtype=torch.half, device='cuda')
  out.fill_(float('nan'))
  if epi == 2:
    out = skip.clone()   # in-place residual, like the engine does
  _lib.check(eng.lib.moe_conv3x3_c64(eng.handle, ctypes.c_void_p(x.data_ptr()), ctypes.c_void_p(out.data_ptr()),
                                     ctypes.c_void_p(out.data_ptr()) if epi == 2 else None,
                                     ctypes.c_void_p(img.data_ptr()), ctypes.c_void_p(bdev.data_ptr()) if bdev is not None else None,
                                     n, h, wd, r, epi, float(param), None))
  torch.cuda.synchronize()
  return out


def conv_suite(eng):
  g = torch.Generator(device='cpu').manual_seed(0)
  cases = [(1, 1, 1, 1, 0), (1, 300, 127, 1, 2), (1, 8, 128, 1, 0), (1, 20, 128, 1, 1), (2, 33, 200, 1, 2), (3, 40, 48, 1, 1), (1, 17, 300, 2, 3), (1, 12, 130, 3, 3),
           (3, 70, 257, 1, 0)]
  for mode_name, simt in (('simt', True), ('tc', False)):
    eng.set_conv_path(simt=simt)
    for (n, h, wd, r, epi) in cases:
      try:
        x = (torch.randn(n, h, wd, 64, generator=g) * 0.5).half().cuda()
        w = (torch.randn(64 * r * r, 64, 3, 3, generator=g) * 0.05).half()
        bias = (torch.randn(64 * r * r, generator=g) * 0.1).half() if epi == 3 else None
        skip = (torch.randn(n, h, wd, 64, generator=g)).half().cuda() if epi == 2 else None
        ref = conv_ref(x, w.cuda(), None if bias is None else bias.cuda(), r, epi, 0.25, skip)
        out = run_conv(eng, x, w, bias, r, epi, 0.25, skip).float()
        d = (out - ref).abs()
        bad = torch.isnan(out).sum().item()
        say('[conv %-20s] n=%d h=%3d w=%3d r=%d epi=%d  max|d|=%.3e mean|d|=%.3e nan=%d ref_max=%.2f' %
            (mode_name, n, h, wd, r, epi, torch.nan_to_num(d, nan=9e9).max().item(), torch.nan_to_num(d).mean().item(), bad, ref.abs().max().item()))
      except Exception as ex:
        say('[conv %-20s] n=%d h=%d w=%d r=%d epi=%d  EXCEPTION %r' % (mode_name, n, h, wd, r, epi, ex))
        if 'launch' in str(ex).lower() or 'cuda' in str(ex).lower():
          return False
  eng.set_conv_path(False)
  return True


def golden_suite(eng):
  sys.path.insert(0, os.path.join(ROOT, 'tests'))
  import helpers as Hh
  for simt in (True, False):
    eng.set_conv_path(simt=simt)
    for name in Hh.case_names():
      try:
        c = Hh.load_case(name)
        y = Hh.run_case_engine(c)
        ref = c['ref']
        d = np.abs(y - ref)
        psnr = 10 * np.log10(1.0 / max(np.mean(d ** 2), 1e-20))
        orc = Hh.run_case_oracle(c, mode='ref16')
        d2 = np.abs(y - orc)
        say('[golden %s] %-16s vs ref fp32: max %.3e psnr %.1f dB | vs oracle ref16: max %.3e  frac>1e-3 %.2e' %
            ('simt' if simt else 'tc  ', name, d.max(), psnr, d2.max(), (d2 > 1e-3).mean()))
      except Exception as ex:
        say('[golden %s] %-16s EXCEPTION %r' % ('simt' if simt else 'tc', name, ex))
        traceback.print_exc()
  eng.set_conv_path(False)


def timing(eng):
  sys.path.insert(0, os.path.join(ROOT, 'tests'))
  import helpers as Hh
  for key, scale, shape, no_pt in (('a2', 2, (3, 1080, 1920), False), ('lite2', 2, (3, 1080, 1920), False), ('lite4', 4, (3, 1080, 1920), False),
                                   ('a3', 3, (3, 1080, 1920), False), ('dn_lite15', 1, (3, 1080, 1920), False), ('a4', 4, (3, 2160, 3840), False)):
    try:
      eng.set_conv_path(simt=False, no_fuse=no_pt)
      sd = Hh.load_weights(key)
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
      t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
      eng.profile(True)
      t0.record()
      for _ in range(3):
        y = run(x)
      t1.record(); torch.cuda.synchronize()
      eng.profile(False)
      pr = eng.profile_read()
      say('[prof] %s' % key, {k: '%.2f ms/frame, %d launches, %.0f T(FLOP|B)/s' % (v[0] / 3, v[2] // 3, v[1] / max(v[0], 1e-9) / 1e9) for k, v in pr.items()},
          'conv TFLOP/s %.0f' % (pr['conv3x3'][1] / (pr['conv3x3'][0] * 1e-3) / 1e12), 'head GB/s %.0f' % (pr['head'][1] / (pr['head'][0] * 1e-3) / 1e9))
      ms = t0.elapsed_time(t1) / 3
      say('[time] %s %s tiles=%d  %.2f ms/frame  %.1f MPix/s out' % (key, shape, len(opt.plan.tiles), ms, y.shape[1] * y.shape[2] / ms / 1e3))
      del y, x, opt
      torch.cuda.empty_cache()
    except Exception as ex:
      say('[time] %s EXCEPTION %r' % (key, ex))
      traceback.print_exc()


def main():
  say('device', torch.cuda.get_device_name(0), torch.cuda.get_device_capability(0))
  eng = IP.getEngine(0)
  ok = conv_suite(eng)
  if ok:
    golden_suite(eng)
    timing(eng)
  say('launches', eng.launches())


if __name__ == '__main__':
  main()
