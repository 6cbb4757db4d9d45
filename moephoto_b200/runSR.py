"""Engine-backed stand-in for MoePhoto's python/runSR.py (reference: runSR.py:9-49).

Public names and meanings are the reference's — `ramCoef`, `mode_switch`, `sr`, `getOpt` and the fields of the
returned Option — because procedure.py:13-14,67,72,171 reads exactly those.  Served here: the `a*` / `p*` nets
(models.Net2x/3x/4x) and MoeNet_lite2 (`lite2/4/8`); the `gan*` rows (RRDBNet) are another model family and stay on
the stock code (install.py).  Like the reference (runSR.py:35-36), getOpt answers None for a model+scale it has no
row for.
"""
import numpy as np

from . import imageProcess as _ip
from .config import config
from .models import LiteNet, Net2x, Net3x, Net4x

# Bytes of working memory per input pixel-plane that the reference calibrated, columns = (CPU fp32, GPU fp32,
# GPU fp16) — rows 0,1,2,4,6,7 of the table at runSR.py:9.  Kept to the digit: the numbers choose the tile grid and the
# tile grid is part of the numerics.
_BYTES_PER_PIXEL = {
  'Net2x': (10888.4, 4971.7, 2473.), 'Net3x': (24248., 8253.9, 6120.), 'Net4x': (41951.3, 16788.7, 7029.7),
  'lite2': (3678., 4712.1, 3223.2), 'lite4': (10803., 10944., 5880.5), 'lite8': (40915., 50049., 27899),
}
ramCoef = .9 / np.array([_BYTES_PER_PIXEL[k] for k in ('Net2x', 'Net3x', 'Net4x', 'lite2', 'lite4', 'lite8')])


def _registry():
  rows = {}
  for family in 'ap':                                    # runSR.py:11-16
    for i, ctor in enumerate((Net2x, Net3x, Net4x)):
      rows['%s%d' % (family, i + 2)] = ('./model/%s%d/model_new.pth' % (family, i + 2), ctor, ramCoef[i])
  for i, (scale, ckpt) in enumerate(((2, 'model.pth'), (4, 'model_4.pth'), (8, 'model_8.pth'))):   # runSR.py:21-23
    rows['lite%d' % scale] = ('./model/lite/' + ckpt, (lambda s: (lambda: LiteNet(upscale=s)))(scale), ramCoef[3 + i])
  return rows


mode_switch = _registry()   # name -> (checkpoint, constructor, ramCoef row), the reference's row shape


def sr(opt):
  """the closure procedure.py:72 composes: x (C,H,W) -> (C,s*H,s*W); with test-time ensemble the mean of the passes"""
  run = _ip.ensemble(opt)
  if not opt.ensemble:
    return run
  return lambda x: run(x) / (opt.ensemble + 1)


def getOpt(optSR, weights=None):
  """optSR = {'model': 'a'|'p'|'lite', 'scale': int, 'ensemble'?: 0..7}.  `weights` (a state dict) replaces the
  checkpoint file — tests and benchmarks run without a MoePhoto tree around them."""
  row = mode_switch.get('{}{}'.format(optSR['model'], optSR['scale']))
  if row is None:
    return None
  checkpoint, constructor, coef = row
  wanted = optSR.get('ensemble', None)
  opt = _ip.Option(checkpoint)
  vars(opt).update(
    mode=optSR['model'], scale=optSR['scale'], modelDef=constructor,
    fixChannel=0,                                          # planes are a batch: the budget divides by the plane count
    squeeze=lambda x: x.squeeze(1), unsqueeze=lambda x: x.unsqueeze(1),
    padding=9 if optSR['scale'] == 3 else 5,               # runSR.py:41
    ensemble=wanted if wanted is not None and 0 <= wanted <= 7 else config.ensembleSR,
    ramCoef=coef[config.getRunType()],
    cropsize=config.getConfig()[0],
  )
  cache_key = None if weights is not None else 'SR{}{}'.format(optSR['model'], optSR['scale'])
  opt.modelCached = _ip.initModel(opt, checkpoint if weights is None else weights, cache_key)
  return opt
