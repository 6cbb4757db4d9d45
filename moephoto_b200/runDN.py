"""Drop-in for MoePhoto's python/runDN.py on the dn_lite models (runDN.py:9-38): same `ramCoef`,
`mode_switch` row shape (checkpoint, constructor, ramCoef, squeezeDim, padding, align) and `getOpt`.
Unknown model names raise KeyError like the reference; the other denoisers (SEDN, MPRNet, NAFNet,
ImageCleaning) are other model families and not on this path.
"""
import numpy as np
from .imageProcess import initModel, Option
from .models import NetDN
from .config import config

ramCoef = .95 / np.array([[2700., 2400., 1253.4]])   # runDN.py:9 row 0
mode_switch = {
  'lite5': ('./model/dn_lite5/model_new.pth', NetDN, ramCoef[0], 1, 7, 8),
  'lite10': ('./model/dn_lite10/model_new.pth', NetDN, ramCoef[0], 1, 7, 8),
  'lite15': ('./model/dn_lite15/model_new.pth', NetDN, ramCoef[0], 1, 7, 8),
}


def getOpt(optDN, weights=None):
  model = optDN['model']
  opt = Option(mode_switch[model][0])
  _, opt.modelDef, ramCoef, sd, opt.padding, opt.align = mode_switch[model]
  opt.strength = optDN.get('strength', 1.0)
  opt.ramCoef = ramCoef[config.getRunType()]
  opt.cropsize = config.getConfig()[1 if model[:4] == 'lite' else 2]
  opt.modelCached = initModel(opt, weights if weights is not None else opt.model, None if weights is not None else 'DN' + model)
  if sd:
    opt.fixChannel = 0
    opt.squeeze = lambda x: x.squeeze(sd)
    opt.unsqueeze = lambda x: x.unsqueeze(sd)
  return opt
