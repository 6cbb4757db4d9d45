"""Architecture descriptors standing in for the reference's nn.Module constructors on this path
(models.py:125-164).  `mode_switch` tables keep the reference's (checkpoint, constructor, ramCoef...)
row shape; here the "constructor" only names the network family the engine implements natively.
"""
from . import weights as _w


class _Arch:
  arch = 0
  filters = 64
  scale = 1


class Net2x(_Arch):      # models.py:125-133
  arch, scale = _w.ARCH_NET2X, 2


class Net3x(_Arch):      # models.py:135-143
  arch, scale = _w.ARCH_NET3X, 3


class Net4x(_Arch):      # models.py:145-154
  arch, scale = _w.ARCH_NET4X, 4


class NetDN(_Arch):      # models.py:158-164
  arch, filters, scale = _w.ARCH_NETDN, 48, 1


class LiteNet(_Arch):    # MoeNet_lite2.Net (MoeNet_lite2.py:22-54); `upscale` = 2, 4 or 8
  arch, filters = _w.ARCH_LITE, 48

  def __init__(self, upscale=2):
    self.scale = upscale
