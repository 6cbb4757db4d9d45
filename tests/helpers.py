"""Shared helpers of the test-suite: golden fixtures (tests/golden, generated from the unmodified
reference by make_golden.py), the oracle runners, and the engine runners that go through the
reference-shaped API (runSR.getOpt / runSR.sr / runDN.getOpt / imageProcess.RGBFilter)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
  sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, 'tests', 'golden')
_cases = None
_bands = None
WEIGHT_OF = {('sr', 2): 'a2', ('sr', 3): 'a3', ('sr', 4): 'a4', ('dn', 'lite15'): 'dn_lite15', ('dn', 'lite5'): 'dn_lite5'}


def _load():
  global _cases
  if _cases is None:
    z = np.load(os.path.join(GOLD, 'cases.npz'))
    meta = json.loads(bytes(z['meta']).decode())
    _cases = (z, meta)
  return _cases


def case_names():
  return list(_load()[1].keys())


def load_case(name):
  z, meta = _load()
  c = dict(meta[name])
  c['name'] = name
  c['img'] = z[name + '.img']
  c['ref'] = z[name + '.ref']                                  # the reference's fp32 CPU output
  c['ref16'] = z[name + '.ref16'].astype(np.float32)           # the reference in its GPU fp16 configuration (executed on CPU)
  c['alpha'] = z[name + '.alpha'] if (name + '.alpha') in z.files else None
  c['weights'] = _weights_key(c)
  return c


def _weights_key(c):
  if c.get('model') == 'lite':
    return 'lite%d' % c['arg']
  if c.get('model') == 'p':
    return 'p%d' % c['arg']
  return WEIGHT_OF[(c['kind'], c['arg'])]


def band_case_names():
  return list(_load_bands()[1].keys())


def _load_bands():
  global _bands
  if _bands is None:
    z = np.load(os.path.join(GOLD, 'bands.npz'))
    _bands = (z, json.loads(bytes(z['meta']).decode()))
  return _bands


def load_band_case(name):
  """large golden: the image, the tile list and `windows` = [((y0,y1,x0,x1), fp16-configuration reference output there)]"""
  z, meta = _load_bands()
  c = dict(meta[name])
  c['name'] = name
  c['img'] = z[name + '.img']
  c['alpha'] = None
  c['ensemble'] = 0
  c['weights'] = _weights_key(c)
  c['windows'] = [(tuple(w), z['%s.win%d' % (name, i)].astype(np.float32)) for i, w in enumerate(c['windows'])]
  return c


def load_weights(key):
  """fp16 copy of a reference checkpoint (tests/golden/weights_<key>.npz) as a float32 numpy state dict"""
  return {k: v.astype(np.float32) for k, v in np.load(os.path.join(GOLD, 'weights_%s.npz' % key)).items()}


def case_input(c, dtype=np.float32):
  from oracle import tiling as T
  x = T.to_planar(c['img'], 8, dtype)
  if c['alpha'] is not None:
    x = np.concatenate([x, c['alpha'][None].astype(dtype)], 0)
  return x


def oracle_plan(c, planes=3, transposed=False):
  from oracle import tiling as T
  h, w = c['img'].shape[:2]
  if transposed:
    h, w = w, h
  return T.make_plan((planes, h, w), c['ram'], c.get('ram_coef', c.get('ram_coef_gpu')), c['pad'], c['scale'], 8, c['crop'])


def run_case_oracle(c, mode='fp32', backend='c'):
  from oracle import net as N, tiling as T
  sd = load_weights(c['weights'])
  dt = np.float32 if mode == 'fp32' else np.float16
  x = case_input(c, dt).astype(np.float32)
  plan = oracle_plan(c)
  net = lambda a: N.forward(sd, a, mode=mode, backend=backend)
  if c['kind'] == 'sr':
    if c.get('ensemble', 0):
      return T.ensemble(net, x, plan, oracle_plan(c, transposed=True), c['ensemble'], dt).astype(np.float32)
    return T.do_crop(net, x, plan, dt).astype(np.float32)
  return T.rgb_filter(net, x, plan, 1.0, dt).astype(np.float32)


def run_case_engine(c, cpu_bias=False):
  """through the reference-shaped API, on the GPU.  cpu_bias: biased convolutions round as in the CPU execution of the
  half model that produced the `.ref16` goldens (engine switch bias_fused); default = as on the reference's GPU path"""
  import torch
  from moephoto_b200 import runSR, runDN, imageProcess as IP
  from moephoto_b200.config import config
  sd = load_weights(c['weights'])
  config.freeMemOverride = c['ram']
  if cpu_bias:
    IP.getEngine().set_conv_path(bias_fused=True)
  x = IP.toTorch(8)(c['img'])
  if c['alpha'] is not None:
    x = torch.cat([x, torch.from_numpy(c['alpha'])[None].to(x.device, x.dtype)], 0)
  try:
    if c['kind'] == 'sr':
      config.crop_sr = c['crop'] if c['crop'] else 'auto'
      opt = runSR.getOpt({'model': c.get('model', 'a'), 'scale': c['arg'], 'ensemble': c.get('ensemble', 0)}, weights=sd)
      y = runSR.sr(opt)(x)
    else:
      config.crop_dn = c['crop'] if c['crop'] else 'auto'
      opt = runDN.getOpt({'model': c['arg']}, weights=sd)
      y = IP.RGBFilter(opt)(x)
    assert [list(t) for t in opt.plan.tiles] == c['tiles'], 'engine tile plan differs from the reference plan'
    return y.float().cpu().numpy()
  finally:
    config.freeMemOverride = None
    config.crop_sr = config.crop_dn = 'auto'
    if cpu_bias:
      IP.getEngine().set_conv_path()


def psnr(a, b):
  return 10 * np.log10(1.0 / max(float(np.mean((a.astype(np.float64) - b) ** 2)), 1e-20))


def is_white_noise(c):
  return c['name'].endswith('_rand')


def has_frm(c):
  return c.get('model') == 'lite'


def assert_ref16_bar(y, c, ref=None, what='output'):
  """THE parity bar against the reference's fp16-configuration output (tests/test_oracle_golden.py docstring):
  smooth images max-abs <= 1e-3 and PSNR >= 75 dB; uniform white noise max-abs <= 2e-3, <= 0.2 % of the pixels beyond 1e-3,
  PSNR >= 68 dB; MoeNet_lite2 (FRM gates) one ulp more on the maximum.  Returns (max, mean, PSNR) for reporting."""
  ref = c['ref16'] if ref is None else ref
  assert y.shape == ref.shape
  d = np.abs(y - ref)
  p = psnr(y, ref)
  if is_white_noise(c) and has_frm(c):
    # MoeNet_lite2 on white noise: a flipped fp16 rounding of one FRM gate (48 per block and plane) rescales a whole channel
    # of the whole tile (models.py:287) — the reference's own cuDNN arithmetic is as far from its CPU arithmetic
    # (test_reference_gpu_port_*, profiles/r02_parity_gpu.txt)
    assert d.max() <= 4e-3 and (d > 1e-3).mean() <= 1e-2 and p >= 66.0, (c['name'], what, d.max(), (d > 1e-3).mean(), p)
  elif is_white_noise(c):
    assert d.max() <= 2e-3 and (d > 1e-3).mean() <= 2e-3 and p >= 68.0, (c['name'], what, d.max(), (d > 1e-3).mean(), p)
  elif has_frm(c):
    assert d.max() <= 2e-3 and (d > 1e-3).mean() <= 1e-4 and p >= 75.0, (c['name'], what, d.max(), p)   # a gate flip: two ulps on isolated pixels
  else:
    assert d.max() <= 1e-3 and p >= 75.0, (c['name'], what, d.max(), p)
  return float(d.max()), float(d.mean()), p


def assert_cross_platform_bar(y, c, ref=None, what='output'):
  """The engine in its DEFAULT numerics (biased convolutions round twice, as aten + cuDNN do on the GPU) against a golden
  the reference's CPU execution produced (bias inside the convolution): one rounding differs per biased convolution
  (a2/a3: 1 per branch, a4: 2, lite: 1-3 + the FRM gate), so the two are 1 ulp apart on ~28 % of those layers' outputs.
  Measured for the ORACLE's two modes against each other (profiles/r02_parity_oracle_vs_reference.txt): 76-77 dB smooth,
  66-70 dB white noise.  Bar: max-abs <= 2e-3 (smooth) / 4e-3 (white noise), PSNR >= 70 / 64 dB."""
  ref = c['ref16'] if ref is None else ref
  d = np.abs(y - ref)
  p = psnr(y, ref)
  if is_white_noise(c):
    assert d.max() <= 4e-3 and p >= 64.0, (c['name'], what, d.max(), p)
  else:
    assert d.max() <= 2e-3 and p >= 70.0, (c['name'], what, d.max(), p)
  return float(d.max()), float(d.mean()), p
