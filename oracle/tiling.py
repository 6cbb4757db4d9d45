"""TEST INFRASTRUCTURE — CPU oracle for the tiling / stitching runtime.  Not product code.

Restates in numpy (integer logic in plain Python) what /root/reference/python/imageProcess.py does
around the network:
  ceilBy / alignF            imageProcess.py:552-556
  getAnchors                 imageProcess.py:19-35
  getPad (reflect-then-zero) imageProcess.py:48-56
  solveRam (scalar k branch) imageProcess.py:61-71
  prepare                    imageProcess.py:73-118
  blend                      imageProcess.py:120-131
  doCrop                     imageProcess.py:157-172
  _RGBFilter / strengthOp    imageProcess.py:370-377, :562
  ensemble (+ runSR.sr)      imageProcess.py:558-572, runSR.py:26
  toTorch / toFloat+toOutput imageProcess.py:259-263, :238-257
Pinned against the unmodified reference by tests/test_oracle_vs_reference.py (tile lists compared
tuple-for-tuple over a sweep of shapes, stitched output compared bit-for-bit with an identity "net").
"""
import math
import numpy as np

MIN_SIZE = 28   # imageProcess.py:553


def ceil_to(v, d):
  """ceilBy(d)(v) for power-of-two d (imageProcess.py:551); d == 1 is the identity."""
  v = int(v)
  return v if d == 1 else -(-v // d) * d


def anchors(size, span, tile, pad, align, sc):
  """imageProcess.py:19-35.  Returns (starts, ends, clip, steps, ends_scaled)."""
  body = tile - 2 * pad
  steps = 1 if tile >= ceil_to(size, align) else max(2, int(math.ceil(span / body)))
  starts = [0] + [k * body + pad for k in range(1, steps)]
  ends = [s + tile for s in starts]
  ends_sc = [e * sc for e in ends]
  if steps > 1:
    starts[-1] = size - ceil_to(size - ends[-2] + pad, align)
    ends[-1] = size
    clip = int((ends[-2] - size) * sc)
  else:
    ends[-1] = ceil_to(size, align)
    clip = 0
  ends_sc[-1] = size * sc
  return starts, ends, clip, steps, [int(e) for e in ends_sc]


def pixel_budget(ram, channels, ram_coef, planes):
  """solveRam with a scalar coefficient (imageProcess.py:61-63, called at :75):
  n = ram / (fixChannel or c) * (ramCoef / shape[0])."""
  return ram / channels * (ram_coef / planes if planes else 1.)


class Plan:
  """Everything prepare() returns (imageProcess.py:118) in explicit form."""
  def __init__(self):
    self.tiles = []          # (top, bottom, left, right, topT, leftT, bsc, rsc) — iterClip() order
    self.pad_h = 0           # rows appended at the bottom (reflect first, zeros if size is tiny)
    self.pad_w = 0
    self.out_h = self.out_w = 0
    self.pad_sc = 0
    self.scale = 1
    self.tile_h = self.tile_w = 0


def make_plan(shape, ram, ram_coef, pad, sc, align=8, cropsize=0, fix_channel=0):
  """imageProcess.py:73-118.  `shape` is the tensor shape doCrop saw, (C,H,W)."""
  c, h, w = shape[-3], shape[-2], shape[-1]
  n = pixel_budget(ram, fix_channel or c, ram_coef, shape[0])
  s = ceil_to(MIN_SIZE + pad * 2, align)
  if n < s * s:
    raise MemoryError('Free memory space is {} bytes, which is not enough.'.format(ram))
  ph, pw = max(1, h - pad * 3), max(1, w - pad * 3)
  # candidate tile heights (multiples of align) and, for each, the widest width the budget allows
  ns = np.arange(s / align, int(n / (align * s)) + 1, dtype=int)
  ms = (n / (align * align) / ns).astype(int)
  ns, ms = ns * align, ms * align
  rows = np.ceil(ph / (ns - 2 * pad)).clip(2)
  cols = np.ceil(pw / (ms - 2 * pad)).clip(2)
  rows[ns >= h] = 1
  cols[ms >= w] = 1
  count = rows * cols
  best = np.argwhere(count == count.min()).squeeze(1)
  pick = best[np.abs(best - len(count) / 2).argmin()]   # fewest tiles, then closest to square
  ah, aw, acs = ceil_to(h, align), ceil_to(w, align), ceil_to(cropsize, align)
  ih, iw = int(ns[pick]), int(ms[pick])
  if cropsize > 0:
    ih, iw = min(acs, ih), min(acs, iw)
  ih, iw = min(ah, ih), min(aw, iw)
  sh, eh, clip_h, step_h, bh = anchors(h, ph, ih, pad, align, sc)
  sw, ew, clip_w, step_w, bw = anchors(w, pw, iw, pad, align, sc)
  p = Plan()
  p.scale, p.pad_sc, p.out_h, p.out_w = sc, int(pad * sc), int(h * sc), int(w * sc)
  p.tile_h, p.tile_w = ih, iw
  # an axis covered by ONE tile is padded up to the alignment (getPad, :100-110); a multi-tile axis is not
  p.pad_h = ah - h if step_h == 1 else 0
  p.pad_w = aw - w if step_w == 1 else 0
  for i in range(step_h):
    top_t = clip_h if i == step_h - 1 else (0 if i == 0 else p.pad_sc)
    for j in range(step_w):
      left_t = clip_w if j == step_w - 1 else (0 if j == 0 else p.pad_sc)
      p.tiles.append((sh[i], eh[i], sw[j], ew[j], top_t, left_t, bh[i], bw[j]))
  return p


def blend_ramp(pad_sc, dtype=np.float32):
  """imageProcess.py:109: sigmoid((arange(padSc)/padSc - .5) * 9), computed in the canvas dtype."""
  t = np.arange(pad_sc).astype(dtype)
  t = ((t / dtype(pad_sc)).astype(dtype) - dtype(.5)).astype(dtype)
  t = (t * dtype(9)).astype(dtype)
  return (1.0 / (1.0 + np.exp(-t.astype(np.float32)))).astype(dtype)


def _pad_axis(x, axis, extra):
  """reflect by min(size-1, extra), then zeros (imageProcess.py:48-56)."""
  if extra <= 0:
    return x
  size = x.shape[axis]
  refl = max(0, min(size - 1, extra))
  widths = [(0, 0)] * x.ndim
  if refl:
    widths[axis] = (0, refl)
    x = np.pad(x, widths, mode='reflect')
  if extra - refl:
    widths[axis] = (0, extra - refl)
    x = np.pad(x, widths, mode='constant')
  return x


def _blend_axis(r, x, lt, pad, axis, ramp, q):
  """imageProcess.py:120-131.  r: new tile, x: what the canvas holds there.  Returns (kept r, kept x)."""
  l = r.shape[axis]
  if lt < 0:
    lt += l
  if lt < 1:
    return r, x
  start = lt - pad
  sl = [slice(None)] * r.ndim
  sl[axis] = slice(start, lt)
  b, bx = r[tuple(sl)], x[tuple(sl)]
  shape = [1] * r.ndim
  shape[axis] = pad
  wgt = ramp.reshape(shape)
  b = q(bx + q(wgt * q(b - bx)))
  sl[axis] = slice(lt, None)
  tail = r[tuple(sl)]
  sl[axis] = slice(start, None)
  return np.concatenate([b, tail], axis), x[tuple(sl)]


def do_crop(net, x, plan, dtype=np.float32, ramp=None):
  """imageProcess.py:157-172.  x: (C,H,W); net maps (C,1,h,w)->(C,1,s*h,s*w).  The canvas is `dtype`
  (float32 = reference CPU path; float16 = reference GPU path, every elementwise op rounds)."""
  q = (lambda a: a.astype(dtype)) if dtype != np.float32 else (lambda a: a)
  if ramp is None:   # (a caller may pass the reference's own ramp: libm sigmoids differ in the last ulp)
    ramp = blend_ramp(plan.pad_sc, dtype) if plan.pad_sc else np.zeros(0, dtype)
  ramp = np.asarray(ramp, dtype=dtype)
  sc, psc = plan.scale, plan.pad_sc
  xp = _pad_axis(_pad_axis(x, -1, plan.pad_w), -2, plan.pad_h)[:, None]
  canvas = np.zeros((x.shape[0], plan.out_h, plan.out_w), dtype=dtype)   # reference: new_empty
  for top, bottom, left, right, top_t, left_t, bsc, rsc in plan.tiles:
    r = np.asarray(net(np.ascontiguousarray(xp[..., top:bottom, left:right])))[:, 0].astype(dtype)
    r = r[..., :bsc - top * sc, :rsc - left * sc]                         # unpad (:100-108)
    t = canvas[..., top * sc:bsc, left * sc:rsc]
    r1, t1 = _blend_axis(r, t, top_t, psc, -2, ramp, q)
    r2, _ = _blend_axis(r1, t1, left_t, psc, -1, ramp, q)
    hh, ww = r2.shape[-2:]
    canvas[..., bsc - hh:bsc, rsc - ww:rsc] = r2
  return canvas


# the seven dihedral passes of the test-time ensemble (imageProcess.py:564-568): pass k applies FWD[k] to the image,
# doCrop on the plan of that orientation (`which`: the transposed plan where the pass swaps H and W), then INV[k]
_t = lambda a: np.swapaxes(a, -1, -2)
_fx = lambda a: a[..., ::-1]
_fxy = lambda a: a[..., ::-1, ::-1]
_seq = lambda *fs: (lambda a: [a := f(a) for f in fs][-1])
ENSEMBLE_FWD = [_t, _fx, _fxy, _seq(_fx, _t), _seq(_t, _fx), _seq(_t, _fx, _t), _seq(_fxy, _t)]
ENSEMBLE_INV = [_t, _fx, _fxy, ENSEMBLE_FWD[4], ENSEMBLE_FWD[3], ENSEMBLE_FWD[5], ENSEMBLE_FWD[6]]
ENSEMBLE_TRANSPOSED = [True, False, False, True, True, False, True]


def ensemble(net, x, plan, plan_t, k, dtype=np.float32, ramp=None):
  """runSR.sr (runSR.py:26) over imageProcess.ensemble (:569-572): (doCrop(x) + sum of the first k dihedral passes)
  / (k + 1), every add and the division rounded to `dtype`.  plan_t: the plan of the transposed image (:142-152)."""
  q = (lambda a: a.astype(dtype)) if dtype != np.float32 else (lambda a: a)
  total = do_crop(net, x, plan, dtype, ramp)
  for fwd, inv, tr in list(zip(ENSEMBLE_FWD, ENSEMBLE_INV, ENSEMBLE_TRANSPOSED))[:k]:
    y = do_crop(net, np.ascontiguousarray(fwd(x)), plan_t if tr else plan, dtype, ramp)
    total = q(total.astype(np.float32) + inv(y).astype(np.float32))
  return q(total.astype(np.float32) / np.float32(k + 1)) if k else total


def rgb_filter(net, img, plan, strength=1.0, dtype=np.float32):
  """imageProcess.py:370-377 + strengthOp :562.  img (3|4,H,W): alpha bypasses the filter."""
  alpha = img[3:] if img.shape[0] == 4 else None
  rgb = img[:3] if alpha is not None else img
  y = do_crop(net, rgb, plan, dtype)
  if strength != 1:
    y = (dtype(strength) * y + dtype(1 - strength) * rgb.astype(dtype)).astype(dtype)
  return y if alpha is None else np.concatenate([y, alpha.astype(dtype)], 0)


def to_planar(image, bit_depth, dtype=np.float32):
  """toTorch (imageProcess.py:259-263): HWC uint8 -> CHW /255 ; wider -> float32 /2^bits, then dtype."""
  chw = np.ascontiguousarray(np.transpose(image, (2, 0, 1)))
  if bit_depth <= 8:
    return (chw.astype(np.float32) / np.float32(255)).astype(dtype)
  return (chw.astype(np.float32) / np.float32(1 << bit_depth)).astype(dtype)


def to_output(image, bit_depth):
  """toFloat + toOutput (imageProcess.py:238-257): CHW -> HWC float32, *2^bits, clamp, truncate."""
  quant = 1 << bit_depth
  v = np.transpose(image, (1, 2, 0)).astype(np.float32) * np.float32(quant)
  v = np.clip(v, 0, quant - 1)
  return v.astype(np.uint8 if bit_depth <= 8 else (np.int16 if bit_depth <= 15 else np.int32))
