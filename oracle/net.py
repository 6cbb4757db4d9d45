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

Three arithmetic modes:
  mode='fp32'  : the reference's CPU path (fp32 everywhere).
  mode='ref16' : the reference's GPU path, `model.half()` on half tensors (imageProcess.py:309-318): EVERY aten
                 op rounds its result to IEEE fp16 — conv (fp32 accumulate, one rounding), the bias add that aten
                 issues as a separate op after cudnn_convolution (see _biased), prelu,
                 the ScaleLayer multiply (models.py:73), the residual add (models.py:60), the branch add
                 (models.py:38), and in MoeNet_lite2 the pooled mean, both 1x1 convs of FRM, the sigmoid, the
                 gate multiply and the skip add (models.py:282-287, MoeNet_lite2.py:15-19).  This is the CUDA
                 engine's numerics contract.  Pinned against the UNMODIFIED reference run in
                 half on CPU (tests/test_oracle_vs_reference.py::test_ref16_*; goldens `<case>.ref16`).
  mode='ref16cpu': the same on the CPU, where oneDNN adds the bias inside the convolution (one rounding less per biased
                 conv) — what the committed `.ref16` goldens were produced with; this is the mode pinned bit-close
                 against the unmodified reference, 'ref16' differs from it only in _biased.
  mode='f16io' : round-1 contract, kept for comparison: fp16 storage with ONE rounding per stored tensor
                 (the ARSB tail and the branch sum round once where the reference rounds three times).
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


def _modes(sd, mode):
  """(q, W, S, half): q rounds a stored tensor, W fetches a parameter as the model holds it, S a scalar parameter,
  half = every aten op rounds (mode 'ref16')"""
  if mode not in ('fp32', 'f16io', 'ref16', 'ref16cpu'):
    raise ValueError('unknown oracle mode %r' % (mode,))
  lowp = mode != 'fp32'
  q = _q16 if lowp else (lambda a: a)
  W = (lambda k: _q16(sd[k])) if lowp else (lambda k: sd[k])
  S = lambda k: np.float32(W(k).reshape(-1)[0])
  return q, W, S, mode in ('ref16', 'ref16cpu')


def _biased(conv, q, a, w, b, split):
  """a convolution with a bias in the half model.  On the GPU aten runs cudnn_convolution WITHOUT the bias and then
  output.add_(bias): two ops, two roundings (measured on B200, profiles/r02_cudnn_rounding_probe.log: cuDNN's result
  equals q(q(conv) + bias) up to summation-order flips, 0.13 %, and differs from q(conv + bias) in 28 % of the outputs).
  On the CPU (oneDNN) the bias is added inside the convolution: one rounding."""
  if split:
    return q(q(conv(a, w, None)) + b[None, :, None, None])
  return q(conv(a, w, b))


def forward_lite(sd, x, mode='fp32', backend='c'):
  """MoeNet_lite2.Net.forward (MoeNet_lite2.py:42-54) with LB (:7-20) and FRM (models.py:270-287):
  out = PReLU(conv1x1(x)); t = conv1x1(out); three times t = FRM(conv3x3(PReLU(conv3x3(t)))) + t with
  FRM(v) = v * sigmoid(W1 relu(W0 mean_hw(v) + b0) + b1); res = ures(t), im = uim(out) (each stage:
  PReLU(PixelShuffle2(conv1x1 + bias))); y = conv1x1(res) + conv1x1(im).
  mode 'ref16': every op of the list above rounds to fp16 (the half model on half tensors).
  mode 'f16io': every stored tensor fp16, the FRM mean / gate in fp32 (round-1 contract)."""
  q, W, S, half = _modes(sd, mode)
  qh = q if half else (lambda a: a)                     # roundings only the reference's per-op path has
  split = mode == 'ref16'                               # GPU: the bias add is its own op
  x = np.ascontiguousarray(x, dtype=np.float32)
  out = q(prelu(qh(conv1x1(x, W('conv_input.weight'))), S('relu.weight')))
  t = q(conv1x1(out, W('conv_input2.weight')))
  for name in ('convt_F11', 'convt_F12', 'convt_F13'):
    mid = q(prelu(qh(conv3x3(t, W(name + '.conv_1.weight'), None, backend)), S(name + '.relu.weight')))
    v = q(conv3x3(mid, W(name + '.conv_2.weight'), None, backend))
    m = qh(v.mean(axis=(2, 3), dtype=np.float32))                                                # (N,48) adaptive_avg_pool2d
    fc = lambda a, wk, bk, co, ci: ((q(a @ W(wk).reshape(co, ci).T) + W(bk)) if split else (a @ W(wk).reshape(co, ci).T + W(bk))).astype(np.float32)
    hid = np.maximum(qh(fc(m, name + '.se.conv_du.0.weight', name + '.se.conv_du.0.bias', 3, 48)), 0)
    z = qh(fc(hid, name + '.se.conv_du.2.weight', name + '.se.conv_du.2.bias', 48, 3))
    gate = qh((1.0 / (1.0 + np.exp(-z.astype(np.float32)))).astype(np.float32))
    g4 = gate.astype(np.float32)[:, :, None, None]
    t = q(qh(v * g4) + t) if half else q(v * g4 + t)
  stages = len([k for k in sd if k.startswith('ures.') and k.endswith('.0.weight')])

  def branch(a, name):
    for j in range(stages):
      if half:
        a = _biased(lambda t, w, b: conv1x1(t, w, b), q, a, W('%s.%d.0.weight' % (name, j)), W('%s.%d.0.bias' % (name, j)), split)
      else:
        a = conv1x1(a, W('%s.%d.0.weight' % (name, j)), W('%s.%d.0.bias' % (name, j)))
      a = q(prelu(pixel_shuffle(a, 2), S('%s.%d.2.weight' % (name, j))))
    return a
  y = qh(conv1x1(branch(t, 'ures'), W('convt_R1.weight'))) + qh(conv1x1(branch(out, 'uim'), W('convt_I1.weight')))
  return q(y)


def forward(sd, x, mode='fp32', backend='c'):
  """x: (N,1,h,w) float32 (values already representable in fp16 for the fp16 modes).
  Returns (N,1,s*h,s*w) float32 — the last element of MyNet.forward's list (imageProcess.py:391-395)."""
  arch = arch_of_state_dict(sd)
  if arch == 'lite':
    return forward_lite(sd, x, mode, backend)
  _, ups = ARCH[arch]
  q, W, S, half = _modes(sd, mode)
  qh = q if half else (lambda a: a)
  split = mode == 'ref16'
  cv = lambda a, k, b=None: conv3x3(a, W(k), None if b is None else W(b), backend)

  x = np.ascontiguousarray(x, dtype=np.float32)
  out = q(prelu(qh(cv(x, 'conv_input.weight')), S('relu.weight')))       # models.py:118
  t = q(cv(out, 'conv_input2.weight'))                                    # models.py:119
  for i in range(1, 7):                                                   # models.py:41-43, 76-80
    p = 'convt_F%d.0.' % i
    mid = q(prelu(qh(cv(t, p + 'conv_1.weight')), S(p + 'relu.weight')))
    c2 = cv(mid, p + 'conv_2.weight')
    if half:
      t = q(q(q(c2) * S(p + 'scale.scale')) + t)                          # conv, ScaleLayer (:73), Residual (:60): three ops
    else:
      t = q(t + S(p + 'scale.scale') * c2)

  def branch(a, name):
    for j, r in enumerate(ups):                                           # models.py:29-33
      if half:
        a = _biased(lambda t, w, b: conv3x3(t, w, b, backend), q, a, W('%s.%d.0.weight' % (name, j)), W('%s.%d.0.bias' % (name, j)), split)
      else:
        a = cv(a, '%s.%d.0.weight' % (name, j), '%s.%d.0.bias' % (name, j))
      a = q(prelu(pixel_shuffle(a, r), S('%s.%d.2.weight' % (name, j))))
    hk = ('%s.%d.weight' % (name, len(ups))) if ups else (name + '.weight')
    return qh(cv(a, hk))                                                  # Conv3x3(F,1)

  y = branch(out, 'u') + branch(t, 'convt_R1')                            # models.py:38,121-123
  return q(y)


def forward_torch(sd, x, dtype='float32', device='cpu'):
  """The same forward entirely in PyTorch functional ops (conv2d / pixel_shuffle / prelu / adaptive_avg_pool2d) —
  operation for operation what the reference's nn.Modules execute, without numpy glue:
    dtype='float32', device='cpu'  : the reference's CPU path (oneDNN) — the timed CPU baseline of bench.py;
    dtype='float16', device='cuda' : the reference's GPU path (cuDNN half, every op rounding to fp16) — used on the GPU
                                     box, where /root/reference does not exist, to measure how far the reference's own
                                     GPU arithmetic is from its CPU-executed fp16 goldens (tests/test_gpu_engine.py).
  Cross-checked against forward() in tests/test_oracle_golden.py.  x: (N,1,h,w) ndarray or tensor -> float32 ndarray."""
  import torch
  import torch.nn.functional as F
  dt = getattr(torch, dtype)
  T = lambda k: torch.from_numpy(np.ascontiguousarray(sd[k], dtype=np.float32)).to(device=device, dtype=dt)
  arch = arch_of_state_dict(sd)
  with torch.no_grad():
    x = (torch.as_tensor(np.asarray(x, dtype=np.float32)) if not torch.is_tensor(x) else x).to(device=device, dtype=dt)
    if arch == 'lite':                                                      # MoeNet_lite2.py:42-54
      c1 = lambda a, k, b=None: F.conv2d(a, T(k), None if b is None else T(b))
      c3 = lambda a, k: F.conv2d(a, T(k), None, padding=1)
      out = F.prelu(c1(x, 'conv_input.weight'), T('relu.weight'))
      t = c1(out, 'conv_input2.weight')
      for name in ('convt_F11', 'convt_F12', 'convt_F13'):
        v = c3(F.prelu(c3(t, name + '.conv_1.weight'), T(name + '.relu.weight')), name + '.conv_2.weight')
        y = F.adaptive_avg_pool2d(v, 1)
        y = torch.sigmoid(c1(F.relu(c1(y, name + '.se.conv_du.0.weight', name + '.se.conv_du.0.bias')), name + '.se.conv_du.2.weight', name + '.se.conv_du.2.bias'))
        t = v * y + t
      stages = len([k for k in sd if k.startswith('ures.') and k.endswith('.0.weight')])

      def lbranch(a, name):
        for j in range(stages):
          a = F.prelu(F.pixel_shuffle(c1(a, '%s.%d.0.weight' % (name, j), '%s.%d.0.bias' % (name, j)), 2), T('%s.%d.2.weight' % (name, j)))
        return a
      return (c1(lbranch(t, 'ures'), 'convt_R1.weight') + c1(lbranch(out, 'uim'), 'convt_I1.weight')).float().cpu().numpy()
    _, ups = ARCH[arch]
    cv = lambda a, k, b=None: F.conv2d(a, T(k), None if b is None else T(b), padding=1)
    out = F.prelu(cv(x, 'conv_input.weight'), T('relu.weight'))
    t = cv(out, 'conv_input2.weight')
    for i in range(1, 7):
      p = 'convt_F%d.0.' % i
      t = cv(F.prelu(cv(t, p + 'conv_1.weight'), T(p + 'relu.weight')), p + 'conv_2.weight') * T(p + 'scale.scale') + t

    def branch(a, name):
      for j, r in enumerate(ups):
        a = F.prelu(F.pixel_shuffle(cv(a, '%s.%d.0.weight' % (name, j), '%s.%d.0.bias' % (name, j)), r), T('%s.%d.2.weight' % (name, j)))
      return cv(a, ('%s.%d.weight' % (name, len(ups))) if ups else (name + '.weight'))
    return (branch(out, 'u') + branch(t, 'convt_R1')).float().cpu().numpy()
