"""Engine-backed mirror of the slice of MoePhoto's imageProcess.py that the SR/DN path uses
(SURVEY.md §8a T1-T8, L3, I1-I2).  Same names, argument meaning and error behaviour:

  Option, initModel, prepare, prepareOpt, doCrop, ensemble, RGBFilter, strengthOp,
  toTorch, toFloat, toOutput, toNumPy, toBuffer, getAnchors, modelCache, weightCache

What differs is where the work happens.  The reference's doCrop (imageProcess.py:157-172) is a Python
loop launching ~50 aten kernels per tile and blending with torch ops; here the tile list is computed on
the host exactly as prepare() does (integer logic) and handed to ONE C-ABI call (moe_run_plan) that
pads, convolves, blends and stores every tile on the device.  There is no CPU path.
"""
import ctypes
import math
from copy import copy

import numpy as np
import torch

from . import _lib
from . import weights as _weights
from .config import config

minSize = 28                                   # imageProcess.py:553
identity = lambda x, *_, **__: x
modelCache = {}
weightCache = {}
_engines = {}


def ceilBy(d):                                 # imageProcess.py:551 (d a power of two)
  return lambda x: -(-int(x) // d) * d


alignF = {1: identity}
alignF.update((1 << k, ceilBy(1 << k)) for k in (3, 4, 5, 6, 7, 9))


# ------------------------------------------------------------------------------------------------
# engine / model handles
# ------------------------------------------------------------------------------------------------
class Engine:
  """one MoeEngine per CUDA device (the worker process owns a single device: config.deviceId)"""

  def __init__(self, device_id):
    self.lib = _lib.load()
    h = ctypes.c_void_p()
    _lib.check(self.lib.moe_engine_create(int(device_id), ctypes.byref(h)))
    self.handle, self.device_id = h, int(device_id)
    self.workspace = None

  def launches(self):
    return int(self.lib.moe_engine_launch_count(self.handle))

  def profile(self, enable):
    _lib.check(self.lib.moe_engine_profile(self.handle, int(bool(enable))))

  PROFILE_CLASSES = ('conv_input', 'conv_trunk', 'head', 'conv_up', 'arsb', 'conv_up_head', 'frm', 'reserved')

  def profile_read(self):
    """-> {class: (ms, algorithmic work, launches)} per kernel class (include/moephoto_b200.h) plus the sums 'conv3x3' (every
    tensor-core convolution), 'trunk' (conv_trunk + arsb) and 'up' (conv_up + conv_up_head); resets the counters"""
    n_cls = len(self.PROFILE_CLASSES)
    ms, work, n = (ctypes.c_double * n_cls)(), (ctypes.c_double * n_cls)(), (ctypes.c_int64 * n_cls)()
    _lib.check(self.lib.moe_engine_profile_read(self.handle, ms, work, n))
    d = {k: (ms[i], work[i], n[i]) for i, k in enumerate(self.PROFILE_CLASSES)}
    add = lambda *ks: tuple(sum(d[k][j] for k in ks) for j in range(3))
    d['trunk'], d['up'] = add('conv_trunk', 'arsb'), add('conv_up', 'conv_up_head')
    d['conv3x3'] = add('conv_trunk', 'arsb', 'conv_up', 'conv_up_head')
    return d

  def set_conv_path(self, simt=False, no_pair=False, no_pair_trunk=False, no_fuse=False, static_sched=False, bias_fused=False, no_arsb=False, arsb_smem_mid=False, arsb_solo=False, full_k=False):
    """A/B switches of the engine.  bias_fused: biased convolutions round once, q(conv + bias), as when the reference's half
    model is executed on the CPU (how the `.ref16` goldens were made); default = the GPU's q(q(conv) + bias)."""
    _lib.check(self.lib.moe_engine_set_conv_path(self.handle, int(bool(simt)) | (int(bool(no_pair)) << 1) | (int(bool(no_pair_trunk)) << 2) |
                                                 (int(bool(no_fuse)) << 3) | (int(bool(static_sched)) << 4) | (int(bool(bias_fused)) << 5) | (int(bool(no_arsb)) << 6) | (int(bool(arsb_smem_mid)) << 7) | (int(bool(arsb_solo)) << 8) | (int(bool(full_k)) << 9)))

  def debug_buffer(self, tensor):
    """tensor: int64 CUDA tensor of >= 4 values per SM pair, or None; see moe_engine_debug_buffer"""
    if tensor is None:
      _lib.check(self.lib.moe_engine_debug_buffer(self.handle, None, 0))
    else:
      _lib.check(self.lib.moe_engine_debug_buffer(self.handle, ctypes.c_void_p(tensor.data_ptr()), tensor.numel() * tensor.element_size()))

  def get_workspace(self, nbytes):
    if self.workspace is None or self.workspace.numel() < nbytes:
      self.workspace = None          # let the caching allocator reuse the old block
      self.workspace = torch.empty(int(nbytes), dtype=torch.uint8, device='cuda:%d' % self.device_id)
    return self.workspace


def getEngine(device_id=None):
  device_id = config.deviceId if device_id is None else device_id
  if device_id not in _engines:
    _engines[device_id] = Engine(device_id)
  return _engines[device_id]


def _stream_ptr(device_id):
  return ctypes.c_void_p(torch.cuda.current_stream(device_id).cuda_stream)


class EngineModel:
  """What `opt.modelCached` holds: the network resident on the device (castModel's result in the
  reference, imageProcess.py:311-318).  Callable on a (N,1,h,w) tile like the bare nn.Module."""

  def __init__(self, engine, state_dict):
    self.engine = engine
    self.arch, blob = _weights.pack(state_dict)
    buf = (ctypes.c_uint8 * len(blob)).from_buffer_copy(blob)
    h = ctypes.c_void_p()
    _lib.check(engine.lib.moe_model_load(engine.handle, self.arch, buf, len(blob), ctypes.byref(h)))
    self.handle = h
    self.scale = int(engine.lib.moe_model_scale(h))

  def __del__(self):
    try:
      if getattr(self, 'handle', None):
        self.engine.lib.moe_model_free(self.handle)
        self.handle = None
    except Exception:
      pass

  def __call__(self, x):
    """MyNet.forward on one tile with zero padding at its border: (N,1,h,w) -> [(N,1,s*h,s*w)]"""
    n, _, h, w = x.shape
    plan = TilePlan.single(h, w, self.scale)
    y = run_plan(self, x.reshape(n, h, w), plan)
    return [y.unsqueeze(1)]


def getStateDict(path):                        # imageProcess.py:304-307
  if path not in weightCache:
    weightCache[path] = torch.load(path, map_location='cpu', weights_only=False)
  return weightCache[path]


def initModel(opt, weights=None, key=None, f=None, args=[]):
  """imageProcess.py:319-334.  `weights` is a checkpoint path or a state dict; the model is cached
  process-wide under `key` ('SRa4', 'DNlite15', ...)."""
  if key and key in modelCache:
    return modelCache[key]
  if weights is None:
    raise ValueError('the engine needs the checkpoint weights to build {}'.format(key or opt.model))
  sd = getStateDict(weights) if isinstance(weights, str) else weights
  model = EngineModel(getEngine(), sd)
  want = getattr(getattr(opt, 'modelDef', None), 'arch', None)
  if want is not None and want != model.arch:
    raise ValueError('checkpoint {} is not a {}'.format(opt.model, opt.modelDef.__name__))
  if key:
    modelCache[key] = model
  return model


# ------------------------------------------------------------------------------------------------
# tile plan (host integer logic)
# ------------------------------------------------------------------------------------------------
def getAnchors(s, ns, l, pad, af, sc):
  """imageProcess.py:19-35: tile starts / ends along one axis, the seam anchor of the last tile, the
  tile count and the ends on the canvas."""
  stride = l - 2 * pad
  count = 1 if l >= af(s) else max(2, int(math.ceil(ns / stride)))
  start = [0] + [i * stride + pad for i in range(1, count)]
  end = [a + l for a in start]
  endSc = [int(b * sc) for b in end]
  clip = 0
  if count > 1:
    clip = int((end[-2] - s) * sc)
    start[-1] = s - af(s - end[-2] + pad)
    end[-1] = s
  else:
    end[-1] = af(s)
  endSc[-1] = int(s * sc)
  return start, end, clip, count, endSc


def _tile_size(n, h, w, pad, align, s):
  """imageProcess.py:81-90: among aligned heights ih (and the widest aligned width the pixel budget n
  allows for each) pick the pair giving the fewest tiles; ties -> the middle candidate."""
  ph, pw = max(1, h - pad * 3), max(1, w - pad * 3)
  first, last = int(s / align), int(n / (align * s))
  best, cands = None, []
  for k in range(first, last + 1):
    ih = k * align
    iw = int(n / (align * align) / k) * align
    rows = 1. if ih >= h else max(2., math.ceil(ph / (ih - 2 * pad)))
    cols = 1. if iw >= w else max(2., math.ceil(pw / (iw - 2 * pad)))
    cands.append((rows * cols, ih, iw))
    best = rows * cols if best is None else min(best, rows * cols)
  mid = len(cands) / 2
  pick = min((i for i, c in enumerate(cands) if c[0] == best), key=lambda i: (abs(i - mid), i))
  return cands[pick][1], cands[pick][2], ph, pw


class TilePlan:
  """prepare()'s result in explicit form; `.c` is the MoePlan passed through the C ABI."""

  def __init__(self, tiles, scale, pad_sc, in_h, in_w, pad_h, pad_w, ramp):
    self.tiles = [tuple(int(v) for v in t) for t in tiles]
    self.scale, self.pad_sc = int(scale), int(pad_sc)
    self.in_h, self.in_w, self.pad_h, self.pad_w = int(in_h), int(in_w), int(pad_h), int(pad_w)
    self.out_h, self.out_w = self.in_h * self.scale, self.in_w * self.scale
    self.ramp = np.ascontiguousarray(ramp, dtype=np.float32)
    self._tiles_c = (_lib.MoeTile * len(self.tiles))(*[_lib.MoeTile(*t) for t in self.tiles])
    self.c = _lib.MoePlan(len(self.tiles), self.scale, self.pad_sc, self.in_h, self.in_w, self.pad_h, self.pad_w,
                          self.out_h, self.out_w, self._tiles_c,
                          self.ramp.ctypes.data_as(ctypes.POINTER(ctypes.c_float)))

  @staticmethod
  def single(h, w, scale):
    return TilePlan([(0, h, 0, w, 0, 0, h * scale, w * scale)], scale, 0, h, w, 0, 0, np.zeros(1, np.float32))


def blendRamp(padSc):
  """imageProcess.py:109 evaluated in the canvas dtype (fp16), on the host"""
  if padSc <= 0:
    return np.zeros(1, dtype=np.float32)
  t = ((torch.arange(padSc, dtype=torch.half) / padSc - .5) * 9).float().sigmoid().half()
  return t.float().numpy()


def solveRam(m, c, k):                         # imageProcess.py:61-63 (scalar coefficient branch)
  return m / c * k


def makePlan(shape, ram, opt, pad, sc, align=8, cropsize=0):
  """imageProcess.py:73-118 -> TilePlan"""
  c, h, w = shape[-3], shape[-2], shape[-1]
  n = solveRam(ram, opt.fixChannel or c, float(opt.ramCoef) / shape[0] if shape[0] else 1.)
  af = alignF[align]
  s = af(minSize + pad * 2)
  if n < s * s:
    raise MemoryError('Free memory space is {} bytes, which is not enough.'.format(ram))
  ih, iw, ph, pw = _tile_size(n, h, w, pad, align, s)
  ah, aw, acs = af(h), af(w), af(cropsize)
  if cropsize > 0:
    ih, iw = min(acs, ih), min(acs, iw)
  ih, iw = min(ah, ih), min(aw, iw)
  startH, endH, clipH, stepH, bH = getAnchors(h, ph, ih, pad, af, sc)
  startW, endW, clipW, stepW, wH = getAnchors(w, pw, iw, pad, af, sc)
  padSc = int(pad * sc)
  tiles = []
  for i in range(stepH):
    topT = clipH if i == stepH - 1 else (0 if i == 0 else padSc)
    for j in range(stepW):
      leftT = clipW if j == stepW - 1 else (0 if j == 0 else padSc)
      tiles.append((startH[i], endH[i], startW[j], endW[j], topT, leftT, bH[i], wH[j]))
  # padImage (imageProcess.py:96-108): only an axis covered by a single tile is padded to the alignment
  return TilePlan(tiles, sc, padSc, h, w, ah - h if stepH == 1 else 0, aw - w if stepW == 1 else 0, blendRamp(padSc))


def _padImage(plan):
  """the reference's padImage as a torch callable (kept for API parity; doCrop pads on the device)"""
  def f(x):
    for axis, extra in ((-1, plan.pad_w), (-2, plan.pad_h)):
      if extra > 0:
        size = x.shape[axis]
        refl = max(0, min(size - 1, extra))
        pads = [0, 0, 0, 0]
        if refl:
          pads[1 if axis == -1 else 3] = refl
          x = torch.nn.functional.pad(x, pads, mode='reflect')
        if extra - refl:
          pads = [0, 0, 0, 0]
          pads[1 if axis == -1 else 3] = extra - refl
          x = torch.nn.functional.pad(x, pads)
    return x
  return f


def prepare(shape, ram, opt, pad, sc, align=8, cropsize=0):
  """imageProcess.py:73-118, same 5-tuple: (iterClip, padImage, unpad, outShape, blend ramp)."""
  plan = makePlan(shape, ram, opt, pad, sc, align, cropsize)
  opt.plan = plan
  unpad = lambda im: im[..., :plan.out_h, :plan.out_w]
  b = torch.from_numpy(plan.ramp[:max(plan.pad_sc, 0)]).to(dtype=config.dtype()).view(1, -1)
  return (lambda: iter(plan.tiles)), _padImage(plan), unpad, (*shape[:-2], plan.out_h, plan.out_w), b


def transposeShape(shape):
  t = list(shape)
  t[-1], t[-2] = shape[-2], shape[-1]
  return t


def _plan_is_stale(opt, shape):
  """the reference re-plans on first use, after 29 cached calls, or when the plane count changes (imageProcess.py:136).
  It does NOT re-plan when the image size changes and then stitches with a stale grid; MoePhoto never hits that
  because getOpt makes a fresh Option per request.  Here a size change re-plans instead of failing."""
  if opt.iterClip is None or opt.count > 28 or shape[0] != opt.outShape[0]:
    return True
  return opt.plan is not None and (opt.plan.in_h, opt.plan.in_w) != (shape[-2], shape[-1])


def prepareOpt(opt, shape):
  """imageProcess.py:133-155: make (or reuse) the tile plan cached on `opt`; returns (scale, seam width)."""
  scale, pad = opt.scale, opt.padding
  if not _plan_is_stale(opt, shape):
    opt.count += 1
    return scale, int(pad * scale)
  try:
    free = config.calcFreeMem()
  except Exception:
    raise MemoryError('Can not calculate free memory.')
  opt.count = 0
  opt.outShape = None if (opt.plan is not None and (opt.plan.in_h, opt.plan.in_w) != (shape[-2], shape[-1])) else opt.outShape
  flipped = None
  if opt.ensemble > 0:                         # the dihedral passes that transpose need a plan for the (W,H) image
    flipped = copy(opt)
    flipped.iterClip, flipped.padImage, flipped.unpad, *_ = prepare(transposeShape(shape), free, flipped, pad, scale, opt.align, opt.cropsize)
  opt.iterClip, opt.padImage, opt.unpad, planned, opt.blend = prepare(shape, free, opt, pad, scale, opt.align, opt.cropsize)
  if opt.outShape is None:
    opt.outShape = planned if not opt.oShape else [1, *opt.oShape[1:-2], int(scale * shape[-2]), int(scale * shape[-1])]
  opt.outShape = list(opt.outShape)
  if flipped is not None:
    flipped.blend, flipped.outShape = opt.blend, transposeShape(opt.outShape)
    opt.transposedOpt = flipped
  return scale, int(pad * scale)


# ------------------------------------------------------------------------------------------------
# doCrop
# ------------------------------------------------------------------------------------------------
def run_plan(model, x, plan, out=None, rows=None, workspace=None):
  """x: (planes,H,W) CUDA tensor -> canvas (planes, s*H, s*W) fp16.  `rows` = (lo,hi) canvas row window.  `workspace`: a uint8
  CUDA tensor of at least moe_plan_workspace_bytes — calls that overlap on different streams need one each (default: the
  engine's own, for the single-stream use of the reference's worker)."""
  if not x.is_cuda:
    raise RuntimeError('moephoto_b200 has no CPU path: the input tensor must live on the GPU')
  if x.dtype != torch.half:
    x = x.half()
  if x.stride(-1) != 1 or x.dim() != 3:
    x = x.reshape(-1, x.shape[-2], x.shape[-1]).contiguous()
  planes, h, w = x.shape
  if (h, w) != (plan.in_h, plan.in_w):
    raise ValueError('plan was made for {}x{}, got {}x{}'.format(plan.in_h, plan.in_w, h, w))
  eng = model.engine
  dev = eng.device_id          # the launch stream is the engine's device's (x / out may be peer-mapped memory of another GPU)
  if out is None:
    out = torch.empty((planes, plan.out_h, plan.out_w), dtype=torch.half, device=x.device)
  lo, hi = (0, plan.out_h) if rows is None else rows
  need = eng.lib.moe_plan_workspace_bytes(model.handle, planes, ctypes.byref(plan.c), lo, hi)
  if need == 0:
    _lib.check(_lib.MOE_ERR_INVALID)
  ws = eng.get_workspace(need) if workspace is None else workspace
  if ws.numel() < need:
    raise MemoryError('workspace of {} bytes is too small (need {})'.format(ws.numel(), need))
  _lib.check(eng.lib.moe_run_plan(model.handle, ctypes.c_void_p(x.data_ptr()), x.stride(0), x.stride(1), planes,
                                  ctypes.c_void_p(out.data_ptr()), out.stride(0), out.stride(1),
                                  ctypes.byref(plan.c), lo, hi, ctypes.c_void_p(ws.data_ptr()), ws.numel(),
                                  _stream_ptr(dev)))
  return out


def doCrop(opt, x, *args, **_):
  """imageProcess.py:157-172.  x: (C,H,W) on config.device(); returns the stitched (C,s*H,s*W) tensor."""
  prepareOpt(opt, x.shape)
  opt.outShape[0] = x.size(0)
  out = x.new_empty(opt.outShape, dtype=torch.half) if x.is_cuda else None
  return run_plan(opt.modelCached, x, opt.plan, out).detach()


# dihedral test-time ensemble (imageProcess.py:558-572): pass k of the reference applies trans[k] before doCrop and
# transInv[k] after it; passes whose forward map swaps H and W run on the transposed plan.
_T = lambda x: x.transpose(-1, -2)
_FX = lambda x: x.flip(-1)
_FXY = lambda x: x.flip(-1, -2)
_chain = lambda *fs: (lambda x: _apply_all(fs, x))


def _apply_all(fs, x):
  for f in fs:
    x = f(x)
  return x


# (forward, inverse, needs the transposed plan)
_DIHEDRAL = [
  (_T, _T, True),
  (_FX, _FX, False),
  (_FXY, _FXY, False),
  (_chain(_FX, _T), _chain(_T, _FX), True),
  (_chain(_T, _FX), _chain(_FX, _T), True),
  (_chain(_T, _FX, _T), _chain(_T, _FX, _T), False),
  (_chain(_FXY, _T), _chain(_FXY, _T), True),
]
trans = [d[0] for d in _DIHEDRAL]              # the reference's module-level names
transInv = [d[1] for d in _DIHEDRAL]
getTransposedOpt = lambda opt: opt.transposedOpt
which = [getTransposedOpt if d[2] else identity for d in _DIHEDRAL]
transpose, flip, flip2 = _T, _FX, _FXY


def ensemble(opt):
  """x -> doCrop(x) + sum of the first opt.ensemble dihedral passes (the caller divides, runSR.sr)"""
  def run(x):
    total = doCrop(opt, x)
    for forward, inverse, transposed in _DIHEDRAL[:opt.ensemble]:
      o = opt.transposedOpt if transposed else opt
      total = (total + inverse(doCrop(o, forward(x).contiguous()))).detach()
    return total
  return run


def strengthOp(x, inp, s=1):
  """imageProcess.py:562: x if s == 1 else s*x + (1-s)*inp (in place on x, fp16 rounding per op)"""
  if s == 1:
    return x
  eng = getEngine(x.device.index)
  inp = inp.to(dtype=torch.half).contiguous()
  _lib.check(eng.lib.moe_axpby_f16(eng.handle, ctypes.c_void_p(x.data_ptr()), ctypes.c_void_p(inp.data_ptr()),
                                   float(s), x.numel(), _stream_ptr(x.device.index)))
  return x


def extractAlpha(t):
  """imageProcess.py:345-352: park a 4th (alpha) plane in `t`, hand on the colour planes"""
  def split(im):
    if im.shape[0] != 4:
      return im
    t['im'] = im[3]
    return im[:3]
  return split


def mergeAlpha(t):
  """imageProcess.py:354-363: put the parked alpha plane back behind the filtered colour planes"""
  def join(im):
    if not t:
      return im
    return torch.cat([im, t['im'].to(im.dtype).unsqueeze(0)], 0)
  return join


def _RGBFilter(opt, img):
  """imageProcess.py:370-377: alpha bypasses the network, colour goes through doCrop, then strengthOp"""
  parked = {}
  colour = opt.prepare(extractAlpha(parked)(img))
  filtered = strengthOp(doCrop(opt, colour), colour, opt.strength)
  return mergeAlpha(parked)(filtered)


RGBFilter = lambda opt: lambda img: _RGBFilter(opt, img)


class Option():
  """imageProcess.py:379-395 — the bag of settings getOpt fills and doCrop reads.  Calling it runs the bare network
  on one tile (`modelCached`) and returns the last element of its output list."""
  _DEFAULTS = dict(ramCoef=1e-3, count=0, padding=1, cropsize=0, align=8, fixChannel=1, scale=1, ensemble=0, strength=1.0,
                   outShape=None, oShape=None, iterClip=None, plan=None)

  def __init__(self, path=''):
    vars(self).update(self._DEFAULTS)
    self.model = path
    self.prepare = identity
    self.squeeze = lambda x: x.squeeze(0)
    self.unsqueeze = lambda x: x.unsqueeze(0)

  def __call__(self, x, *args, **kwargs):
    out = self.modelCached(x, *args, **kwargs)
    return out[-1] if isinstance(out, list) else out


# ------------------------------------------------------------------------------------------------
# frame <-> tensor conversions (imageProcess.py:238-263)
# ------------------------------------------------------------------------------------------------
def toNumPy(bitDepth):
  """imageProcess.py:216-229, the video route's first step: (raw bytes, height, width) -> HWC integer frame.  The
  reference converts to float32 on the host here; the engine keeps the integers (toTorch divides on the GPU)."""
  dtype = np.uint8 if bitDepth <= 8 else (np.uint16 if bitDepth <= 16 else np.int32)
  def f(args):
    buffer, height, width = args
    if not buffer:
      return None
    return np.frombuffer(buffer, dtype=dtype).reshape((height, width, 3))
  return f


def toBuffer(bitDepth):
  """imageProcess.py:231-236: HWC integer frame -> raw bytes for the encoder pipe (`tostring` in the reference)"""
  dtype = np.uint8 if bitDepth == 8 else np.uint16
  return lambda im: None if im is None else np.ascontiguousarray(im).astype(dtype, copy=False).tobytes()


def toTorch(bitDepth, dtype=None, device=None, swapRB=False):
  """HWC integer ndarray -> CHW fp16 tensor on the device, /255 (8 bit) or /2^bits; the integer frame
  is what crosses PCIe, the division runs on the GPU."""
  def f(image):
    image = np.ascontiguousarray(image)
    if image.ndim == 2:
      image = image[:, :, None]
    h, w, c = image.shape
    want = np.uint8 if bitDepth <= 8 else np.uint16
    if image.dtype != want:
      image = image.astype(want)
    eng = getEngine()
    dev = 'cuda:%d' % eng.device_id
    raw = torch.from_numpy(image).to(dev, non_blocking=True)
    out = torch.empty((c, h, w), dtype=torch.half, device=dev)
    _lib.check(eng.lib.moe_to_planar_f16(eng.handle, ctypes.c_void_p(raw.data_ptr()), int(bitDepth), h, w, c, int(swapRB),
                                         ctypes.c_void_p(out.data_ptr()), _stream_ptr(eng.device_id)))
    return out
  return f


def toFloat(image):                            # imageProcess.py:238-243 (torch view/cast, kept for API parity)
  image = image.permute(1, 2, 0) if len(image.shape) == 3 else image.squeeze(0)
  return image.to(dtype=torch.float)


def toOutput(bitDepth, swapRB=False, fused=True):
  """CHW fp16 device tensor -> HWC integer ndarray: x2^bits, clamp, truncate (toFloat + toOutput of the
  reference fused into one kernel; only the integer frame crosses PCIe).  An HWC float tensor — what
  the reference's toFloat hands over — is accepted too."""
  quant = 1 << bitDepth
  npdt = np.uint8 if bitDepth <= 8 else np.uint16
  def f(image):
    if image.dtype == torch.float and image.dim() == 3 and image.shape[-1] in (1, 3, 4):
      image = image.permute(2, 0, 1)            # undo toFloat
    image = image.to(dtype=torch.half).contiguous()
    c, h, w = image.shape
    eng = getEngine(image.device.index)
    out = torch.empty((h, w, c), dtype=torch.uint8 if bitDepth <= 8 else torch.int16, device=image.device)
    _lib.check(eng.lib.moe_to_output(eng.handle, ctypes.c_void_p(image.data_ptr()), int(bitDepth), h, w, c, int(swapRB),
                                     ctypes.c_void_p(out.data_ptr()), _stream_ptr(image.device.index)))
    return out.cpu().numpy().view(npdt)
  return f
