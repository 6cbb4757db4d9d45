"""Drop-in for MoePhoto's python/runSR.py on the a*/p* models and MoeNet_lite2 (runSR.py:9-49): same module-level
names `ramCoef`, `mode_switch`, `sr`, `getOpt`, same Option fields — backed by the sm_100a engine.
The `gan*` rows (RRDBNet) are another model family and not on this path; getOpt returns None for
any model+scale it does not know, exactly as the reference does for unknown names (runSR.py:35-36).
"""
import numpy as np
from .imageProcess import ensemble, initModel, Option
from .models import Net2x, Net3x, Net4x, LiteNet
from .config import config

# bytes per input pixel-plane the reference calibrated for (CPU fp32, GPU fp32, GPU fp16), runSR.py:9.
# Kept verbatim: it decides the tile grid, and the tile grid is part of the numerics.
ramCoef = .9 / np.array([[10888.4, 4971.7, 2473.], [24248., 8253.9, 6120.], [41951.3, 16788.7, 7029.7],
                         [3678., 4712.1, 3223.2], [10803., 10944., 5880.5], [40915., 50049., 27899]])   # rows 4, 6, 7 of runSR.py:9
mode_switch = {
  'a2': ('./model/a2/model_new.pth', Net2x, ramCoef[0]),
  'a3': ('./model/a3/model_new.pth', Net3x, ramCoef[1]),
  'a4': ('./model/a4/model_new.pth', Net4x, ramCoef[2]),
  'p2': ('./model/p2/model_new.pth', Net2x, ramCoef[0]),
  'p3': ('./model/p3/model_new.pth', Net3x, ramCoef[1]),
  'p4': ('./model/p4/model_new.pth', Net4x, ramCoef[2]),
  'lite2': ('./model/lite/model.pth', LiteNet, ramCoef[3]),
  'lite4': ('./model/lite/model_4.pth', lambda: LiteNet(upscale=4), ramCoef[4]),
  'lite8': ('./model/lite/model_8.pth', lambda: LiteNet(upscale=8), ramCoef[5]),
}

sr = lambda opt: (lambda x: ensemble(opt)(x) / (opt.ensemble + 1)) if opt.ensemble else ensemble(opt)


def getOpt(optSR, weights=None):
  """optSR: {'model': 'a'|'p', 'scale': 2|3|4, 'ensemble'?: 0..7}.  `weights` (a state dict) overrides
  the checkpoint path — used by tests and benchmarks that have no MoePhoto tree around them."""
  opt = Option()
  opt.mode = optSR['model']
  opt.scale = optSR['scale']
  nmode = opt.mode + str(opt.scale)
  if nmode not in mode_switch:
    return None
  opt.fixChannel = 0
  opt.squeeze = lambda x: x.squeeze(1)
  opt.unsqueeze = lambda x: x.unsqueeze(1)
  opt.padding = 9 if opt.scale == 3 else 5
  opt.model = mode_switch[nmode][0]
  opt.modelDef = mode_switch[nmode][1]
  opt.ensemble = optSR['ensemble'] if 'ensemble' in optSR and (0 <= optSR['ensemble'] <= 7) else config.ensembleSR
  opt.ramCoef = mode_switch[nmode][2][config.getRunType()]
  opt.cropsize = config.getConfig()[0]
  opt.modelCached = initModel(opt, weights if weights is not None else opt.model, None if weights is not None else 'SR' + nmode)
  return opt
