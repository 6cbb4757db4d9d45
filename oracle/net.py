"""TEST INFRASTRUCTURE — CPU oracle for the SR/DN networks.  Not product code: only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import this.

Restates, in numpy + the plain-C primitives of conv_ref.c, the forward pass of
  MyNet.forward        /root/reference/python/models.py:117-123
  multiConvt (eval)    models.py:41-43
  ARSB                 models.py:76-80   (x + scale * conv_2(PReLU(conv_1(x))))
  genUpsampleBlock(r)  models.py:29-33   (PReLU(PixelShuffle_r(conv3x3(x)+bias)))
  Net2x/Net3x/Net4x    models.py:125-154 (u and convt_R1: independent upsample stacks + Conv3x3(64,1))
  NetDN                models.py:158-164 (48 filters, plain Conv3x3(48,1) heads, no input skip)
Weights are addressed by the checkpoint key names produced by initParameters (models.py:16-19):
  conv_input.weight, relu.weight, conv_input2.weight, convt_F{i}.0.conv_{1,2}.weight,
  convt_F{i}.0.relu.weight, convt_F{i}.0.scale.scale, u.*, convt_R1.*

Pinning: tests/test_oracle_vs_reference.py runs this against the unmodified reference (imported in
the build container through oracle/refharness.py) and against tests/golden/*.npz generated from it
(the reference has no golden vectors of its own — SURVEY.md §4 — so parity is pinned by running it).

Two arithmetic modes:
  mode='fp32'  : the reference's CPU path (fp32 everywhere).
  mode='f16io' : the rounding points of the CUDA engine (and, up to two extra roundings in the ARSB
                 tail, of the reference's GPU fp16 path): weights and every stored activation are
                 rounded to IEEE fp16, accumulation stays fp32.  See DESIGN.md §numerics.
Two conv back-ends: 'c' (conv_ref.c, independent of torch) and 'torch' (F.conv2d on CPU = oneDNN,
the arithmetic the reference's CPU path really executes; used for the timed CPU baseline).
"""
import ctypes
import numpy as np

from . import build as _build

_lib = None


def _c():
  global _lib
  if _lib is None:
    lib = ctypes.CDLL(_build.build())
    fp = ctypes.POINTER(ctypes.c_float)
    lib.oracle_conv3x3.argtypes = [fp, fp, fp, fp] + [ctypes.c_int] * 5
    lib.oracle_prelu.argtypes = [fp, ctypes.c_size_t, ctypes.c_float]
    lib.oracle_pixel_shuffle.argtypes = [fp, fp] + [ctypes.c_int] * 5
    _lib = lib
  return _lib


def _ptr(a):
  return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def conv3x3(x, w, b=None, backend='c'):
  """x (N,Cin,H,W) fp32, w (Cout,Cin,3,3), b (Cout,) or None -> (N,Cout,H,W) fp32."""
  x = np.ascontiguousarray(x, dtype=np.float32)
  w = np.ascontiguousarray(w, dtype=np.float32)
  n, cin, h, wd = x.shape
  cout = w.shape[0]
  if backend == 'torch':
    import torch
    import torch.nn.functional as F
    with torch.no_grad():
      y = F.conv2d(torch.from_numpy(x), torch.from_numpy(w),
                   None if b is None else torch.from_numpy(np.ascontiguousarray(b, dtype=np.float32)), padding=1)
    return y.numpy()
  out = np.empty((n, cout, h, wd), dtype=np.float32)
  bb = None if b is None else np.ascontiguousarray(b, dtype=np.float32)
  _c().oracle_conv3x3(_ptr(x), _ptr(w), None if bb is None else _ptr(bb), _ptr(out), n, cin, cout, h, wd)
  return out


def prelu(x, slope):
  return np.where(x >= 0, x, np.float32(slope) * x).astype(np.float32)


def pixel_shuffle(x, r):
  n, c, h, w = x.shape
  co = c // (r * r)
  return x.reshape(n, co, r, r, h, w).transpose(0, 1, 4, 2, 5, 3).reshape(n, co, h * r, w * r)


def _q16(a):
  return a.astype(np.float16).astype(np.float32)


ARCH = {
  # name: (filters, [upsample factors], checkpoint ctor in the reference)
  'net2x': (64, [2]),      # models.py:125-133
  'net3x': (64, [3]),      # models.py:135-143
  'net4x': (64, [2, 2]),   # models.py:145-154
  'netdn': (48, []),       # models.py:158-164
  'lite': (48, None),      # MoeNet_lite2.py:22-54, upscale 2 / 4 / 8 = 1 / 2 / 3 PixelShuffle(2) stages
}


def arch_of_state_dict(sd):
  if 'convt_F11.conv_1.weight' in sd:
    return 'lite'
  f = sd['conv_input.weight'].shape[0]
  if f == 48:
    return 'netdn'
  if 'u.2.weight' in sd:
    return 'net4x'
  return 'net3x' if sd['u.0.0.weight'].shape[0] == 576 else 'net2x'


def to_numpy_state(sd):
  return {k: (v.detach().cpu().numpy() if hasattr(v, 'detach') else np.asarray(v)).astype(np.float32) for k, v in sd.items()}


def conv1x1(x, w, b=None):
  """x (N,Cin,H,W), w (Cout,Cin,1,1) -> (N,Cout,H,W): out[n,co,y,x] = b[co] + sum_ci w[co,ci] in[n,ci,y,x]"""
  y = np.einsum('oc,nchw->nohw', np.asarray(w, dtype=np.float32).reshape(w.shape[0], w.shape[1]), np.asarray(x, dtype=np.float32), optimize=True)
  return (y if b is None else y + np.asarray(b, dtype=np.float32)[None, :, None, None]).astype(np.float32)


def forward_lite(sd, x, mode='fp32', backend='c'):
  """MoeNet_lite2.Net.forward (MoeNet_lite2.py:42-54) with LB (:7-20) and FRM (models.py:270-287):
  out = PReLU(conv1x1(x)); t = conv1x1(out); three times t = FRM(conv3x3(PReLU(conv3x3(t)))) + t with
  FRM(v) = v * sigmoid(W1 relu(W0 mean_hw(v) + b0) + b1); res = ures(t), im = uim(out) (each stage:
  PReLU(PixelShuffle2(conv1x1 + bias))); y = conv1x1(res) + conv1x1(im).
  mode 'f16io' has the CUDA engine's rounding points: every stored tensor fp16, the FRM mean / gate in fp32."""
  q = _q16 if mode == 'f16io' else (lambda a: a)
  W = (lambda k: _q16(sd[k])) if mode == 'f16io' else (lambda k: sd[k])
  S = lambda k: np.float32(W(k).reshape(-1)[0])
  x = np.ascontiguousarray(x, dtype=np.float32)
  out = q(prelu(conv1x1(x, W('conv_input.weight')), S('relu.weight')))
  t = q(conv1x1(out, W('conv_input2.weight')))
  for name in ('convt_F11', 'convt_F12', 'convt_F13'):
    mid = q(prelu(conv3x3(t, W(name + '.conv_1.weight'), None, backend), S(name + '.relu.weight')))
    v = q(conv3x3(mid, W(name + '.conv_2.weight'), None, backend))
    m = v.mean(axis=(2, 3), dtype=np.float32)                                                   # (N,48)
    hid = np.maximum(m @ W(name + '.se.conv_du.0.weight').reshape(3, 48).T + W(name + '.se.conv_du.0.bias'), 0)
    gate = 1.0 / (1.0 + np.exp(-(hid @ W(name + '.se.conv_du.2.weight').reshape(48, 3).T + W(name + '.se.conv_du.2.bias'))))
    t = q(v * gate.astype(np.float32)[:, :, None, None] + t)
  stages = len([k for k in sd if k.startswith('ures.') and k.endswith('.0.weight')])

  def branch(a, name):
    for j in range(stages):
      a = conv1x1(a, W('%s.%d.0.weight' % (name, j)), W('%s.%d.0.bias' % (name, j)))
      a = q(prelu(pixel_shuffle(a, 2), S('%s.%d.2.weight' % (name, j))))
    return a
  y = conv1x1(branch(t, 'ures'), W('convt_R1.weight')) + conv1x1(branch(out, 'uim'), W('convt_I1.weight'))
  return q(y)


def forward(sd, x, mode='fp32', backend='c'):
  """x: (N,1,h,w) float32 (values already representable in fp16 for mode='f16io').
  Returns (N,1,s*h,s*w) float32 — the last element of MyNet.forward's list (imageProcess.py:391-395)."""
  arch = arch_of_state_dict(sd)
  if arch == 'lite':
    return forward_lite(sd, x, mode, backend)
  _, ups = ARCH[arch]
  q = _q16 if mode == 'f16io' else (lambda a: a)
  W = (lambda k: _q16(sd[k])) if mode == 'f16io' else (lambda k: sd[k])
  S = lambda k: np.float32(W(k).reshape(-1)[0])
  cv = lambda a, k, b=None: conv3x3(a, W(k), None if b is None else W(b), backend)

  x = np.ascontiguousarray(x, dtype=np.float32)
  out = q(prelu(cv(x, 'conv_input.weight'), S('relu.weight')))           # models.py:118
  t = q(cv(out, 'conv_input2.weight'))                                    # models.py:119
  for i in range(1, 7):                                                   # models.py:41-43, 76-80
    p = 'convt_F%d.0.' % i
    mid = q(prelu(cv(t, p + 'conv_1.weight'), S(p + 'relu.weight')))
    t = q(t + S(p + 'scale.scale') * cv(mid, p + 'conv_2.weight'))

  def branch(a, name):
    for j, r in enumerate(ups):                                           # models.py:29-33
      a = cv(a, '%s.%d.0.weight' % (name, j), '%s.%d.0.bias' % (name, j))
      a = q(prelu(pixel_shuffle(a, r), S('%s.%d.2.weight' % (name, j))))
    hk = ('%s.%d.weight' % (name, len(ups))) if ups else (name + '.weight')
    return cv(a, hk)                                                      # Conv3x3(F,1)

  y = branch(out, 'u') + branch(t, 'convt_R1')                            # models.py:38,121-123
  return q(y)


def forward_torch(sd, x):
  """The same forward, fp32, entirely in PyTorch CPU functional ops (conv2d / pixel_shuffle / prelu) —
  operation for operation what the reference's nn.Modules execute on its CPU path (oneDNN), without
  numpy glue.  Used for the timed CPU baseline (bench.py) and cross-checked against forward() in
  tests/test_oracle_golden.py.  x: (N,1,h,w) float32 ndarray or tensor -> ndarray."""
  import torch
  import torch.nn.functional as F
  T = lambda k: torch.from_numpy(np.ascontiguousarray(sd[k], dtype=np.float32))
  if arch_of_state_dict(sd) == 'lite':
    raise NotImplementedError('forward_torch covers the Net2x/3x/4x/NetDN baselines only')
  _, ups = ARCH[arch_of_state_dict(sd)]
  with torch.no_grad():
    x = torch.as_tensor(np.asarray(x, dtype=np.float32)) if not torch.is_tensor(x) else x
    cv = lambda a, k, b=None: F.conv2d(a, T(k), None if b is None else T(b), padding=1)
    out = F.prelu(cv(x, 'conv_input.weight'), T('relu.weight'))
    t = cv(out, 'conv_input2.weight')
    for i in range(1, 7):
      p = 'convt_F%d.0.' % i
      t = t + T(p + 'scale.scale') * cv(F.prelu(cv(t, p + 'conv_1.weight'), T(p + 'relu.weight')), p + 'conv_2.weight')

    def branch(a, name):
      for j, r in enumerate(ups):
        a = F.prelu(F.pixel_shuffle(cv(a, '%s.%d.0.weight' % (name, j), '%s.%d.0.bias' % (name, j)), r), T('%s.%d.2.weight' % (name, j)))
      return cv(a, ('%s.%d.weight' % (name, len(ups))) if ups else (name + '.weight'))
    return (branch(out, 'u') + branch(t, 'convt_R1')).numpy()
