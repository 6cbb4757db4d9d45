"""install(): route MoePhoto's a*/p*/lite*/dn_lite* models through the engine inside an existing MoePhoto
process, leaving every other model on the stock code.  Call it once, before `procedure` is imported
(e.g. at the top of python/MoePhoto.py).  It patches three names the step-pipeline builder reads:
  runSR.getOpt / runSR.sr          procedure.py:171, :72
  runDN.getOpt                     procedure.py:171
  imageProcess.RGBFilter           procedure.py:55
An Option produced by the engine carries `.plan`; that is how the patched closures tell the two kinds apart.
"""


def install(moe_config=None):
  import runSR as ref_runSR                 # the stock modules of the MoePhoto tree on sys.path
  import runDN as ref_runDN
  import imageProcess as ref_ip
  from . import runSR as b_runSR, runDN as b_runDN, imageProcess as b_ip
  from .config import config as b_cfg
  if moe_config is None:
    from config import config as moe_config
  if not (moe_config.cuda and moe_config.fp16):
    return False                            # the engine is GPU fp16 only: keep the stock path
  for k in ('deviceId', 'crop_sr', 'crop_dn', 'crop_dns', 'ensembleSR', 'maxGraphicMemoryUsage'):
    if hasattr(moe_config, k):
      setattr(b_cfg, k, getattr(moe_config, k))
  stock_getOptSR, stock_sr, stock_getOptDN, stock_filter = ref_runSR.getOpt, ref_runSR.sr, ref_runDN.getOpt, ref_ip.RGBFilter
  is_engine = lambda opt: getattr(opt, 'plan', 'missing') != 'missing'

  def getOptSR(optSR):
    name = '{}{}'.format(optSR['model'], optSR['scale'])
    return b_runSR.getOpt(optSR) if name in b_runSR.mode_switch else stock_getOptSR(optSR)

  def getOptDN(optDN):
    return b_runDN.getOpt(optDN) if optDN['model'] in b_runDN.mode_switch else stock_getOptDN(optDN)

  ref_runSR.getOpt = getOptSR
  ref_runSR.sr = lambda opt: b_runSR.sr(opt) if is_engine(opt) else stock_sr(opt)
  ref_runDN.getOpt = getOptDN
  ref_ip.RGBFilter = lambda opt: b_ip.RGBFilter(opt) if is_engine(opt) else stock_filter(opt)
  return True
