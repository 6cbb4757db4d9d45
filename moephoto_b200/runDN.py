"""Engine-backed stand-in for MoePhoto's python/runDN.py (reference: runDN.py:9-38) for the dn_lite models.

Same public names — `ramCoef`, `mode_switch` (rows of (checkpoint, constructor, ramCoef, squeezeDim, padding,
align)), `getOpt` — and the same behaviour for an unknown model name (KeyError).  The other denoisers of the
reference's table (SEDN, MPRNet, NAFNet, ImageCleaning) are other model families and stay on the stock code.
"""
import numpy as np

from . import imageProcess as _ip
from .config import config
from .models import NetDN

ramCoef = .95 / np.array([[2700., 2400., 1253.4]])      # row 0 of runDN.py:9: (CPU fp32, GPU fp32, GPU fp16)
mode_switch = {'lite%d' % level: ('./model/dn_lite%d/model_new.pth' % level, NetDN, ramCoef[0], 1, 7, 8)
               for level in (5, 10, 15)}


def getOpt(optDN, weights=None):
  """optDN = {'model': 'lite5'|'lite10'|'lite15', 'strength'?: float}"""
  name = optDN['model']
  checkpoint, constructor, coef, plane_dim, padding, align = mode_switch[name]      # KeyError for unknown names
  opt = _ip.Option(checkpoint)
  vars(opt).update(
    modelDef=constructor, padding=padding, align=align,
    strength=optDN.get('strength', 1.0),
    ramCoef=coef[config.getRunType()],
    cropsize=config.getConfig()[1 if name.startswith('lite') else 2],
  )
  opt.modelCached = _ip.initModel(opt, checkpoint if weights is None else weights, None if weights is not None else 'DN' + name)
  if plane_dim:                                          # planes are a batch of single-channel images (runDN.py:34-37)
    opt.fixChannel = 0
    opt.squeeze = lambda x: x.squeeze(plane_dim)
    opt.unsqueeze = lambda x: x.unsqueeze(plane_dim)
  return opt
