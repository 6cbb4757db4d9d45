"""The slice of MoePhoto's `config` singleton that the SR/DN path reads (config.py:21-98,
defaultConfig.py): device / dtype / run type, crop sizes, ensemble default and the free-memory probe
that drives the tile size.  When this package is dropped into a MoePhoto tree the host passes its own
`config` object instead — the attribute names are the same.
"""
import torch


class Config:
  def __init__(self):
    self.cuda = True                # defaultConfig.py:10
    self.fp16 = True                # defaultConfig.py:11 — the engine computes in fp16 storage / fp32 accumulate only
    self.deviceId = 0
    self.crop_sr = 'auto'           # defaultConfig.py:3-5
    self.crop_dn = 'auto'
    self.crop_dns = 'auto'
    self.ensembleSR = 0
    self.maxGraphicMemoryUsage = 0
    self.freeMemOverride = None     # tests / benchmarks pin the plan input here

  def getConfig(self):              # config.py:43-44 ('auto' -> 0)
    return tuple(0 if v == 'auto' else v for v in (self.crop_sr, self.crop_dn, self.crop_dns))

  def dtype(self):                  # config.py:73-74
    return torch.half

  def device(self):                 # config.py:76-77
    return torch.device('cuda:{}'.format(self.deviceId))

  def getRunType(self):             # config.py:79-83: 2 = GPU fp16 (index into ramCoef rows)
    return 2

  def getFreeMem(self):             # config.py:50-59; NVML free == cudaMemGetInfo free
    free, _ = torch.cuda.mem_get_info(self.deviceId)
    return free - 2 ** 28

  def calcFreeMem(self, ratio=.9):  # config.py:61-71
    if self.freeMemOverride is not None:
      return int(self.freeMemOverride)
    free = (self.getFreeMem() + torch.cuda.memory_reserved(self.deviceId)) * ratio
    if self.maxGraphicMemoryUsage > 0:
      free = min(free, self.maxGraphicMemoryUsage * 2 ** 20 - torch.cuda.memory_allocated(self.deviceId))
    return int(free)


config = Config()
