"""state dict -> packed weight blob for moe_model_load (layout: csrc/blob.h).

Replaces the tensor side of imageProcess.initModel / castModel (imageProcess.py:311-334): the
reference does `model.load_state_dict(sd); model.to(half, cuda)`; here the same fp16 rounding of every
parameter happens on the host and the 3x3 convolutions are laid out the way the tensor cores read them:
  * 64->64 conv: one 9 x 64 x 128-byte image [tap][out channel][in channel], each 16-byte group of
    input channels XOR-swizzled with (out channel & 7) — the shared-memory image of a K-major
    SWIZZLE_128B operand, so the kernel fetches it with a flat bulk copy;
  * upsample conv 64 -> 64*r*r (+ PixelShuffle(r), models.py:29-33): r*r such images, image (i,j)
    holding output channels c*r*r + i*r + j — the channels PixelShuffle sends to sub-pixel (i,j) — so the
    shuffle is a store address, not a data movement;
  * NetDN's 48 filters are zero-padded to 64 (zero weights keep the padded channels at exactly 0).
Checkpoint key names: models.py:16-19 (initParameters), SURVEY.md §8a N6.
"""
import struct
import numpy as np

ARCH_NETDN, ARCH_NET2X, ARCH_NET3X, ARCH_NET4X, ARCH_LITE = 1, 2, 3, 4, 5
MAGIC, VERSION = 0x42454F4D, 2
SEC_FIRST_W, SEC_SCALARS, SEC_TRUNK_IMG, SEC_UP_IMG, SEC_UP_BIAS, SEC_HEAD_W, SEC_FRM = 1, 2, 3, 4, 5, 6, 7


def _np(v):
  if hasattr(v, 'detach'):
    v = v.detach().cpu().float().numpy()
  return np.asarray(v, dtype=np.float32)


def _h(v):
  """the value the reference's fp16 model holds"""
  return _np(v).astype(np.float16)


def detect_arch(sd):
  if 'convt_F11.conv_1.weight' in sd:
    return ARCH_LITE                      # MoeNet_lite2.Net (MoeNet_lite2.py:22-54)
  f = int(sd['conv_input.weight'].shape[0])
  if f == 48 and 'u.weight' in sd:
    return ARCH_NETDN
  if f != 64:
    raise ValueError('unsupported filter count %d' % f)
  if 'u.2.weight' in sd:
    return ARCH_NET4X
  co = int(sd['u.0.0.weight'].shape[0])
  if co == 576:
    return ARCH_NET3X
  if co == 256:
    return ARCH_NET2X
  raise ValueError('unrecognised upsample width %d' % co)


def conv_image(w16):
  """(Cout<=64, Cin<=64, 3, 3) fp16 -> 73 728-byte swizzled image [tap][row][128 B]."""
  co, ci = w16.shape[:2]
  full = np.zeros((64, 64, 3, 3), dtype=np.float16)
  full[:co, :ci] = w16
  t = full.reshape(64, 64, 9).transpose(2, 0, 1)            # [tap][row=cout][cin]
  t = t.reshape(9, 64, 8, 8)                                 # [tap][row][16-byte group][8 halfs]
  rows = np.arange(64)[:, None]
  src_group = np.arange(8)[None, :] ^ (rows & 7)             # physical group g holds logical group g ^ (row & 7)
  img = t[:, rows, src_group, :]                             # [tap][row][physical group][8]
  return np.ascontiguousarray(img).view(np.uint8).reshape(-1)


def _center(w16):
  """(Cout,Cin,1,1) -> (Cout,Cin,3,3) with the 1x1 filter on the centre tap: a 1x1 convolution run by the 3x3 kernels"""
  full = np.zeros(w16.shape[:2] + (3, 3), dtype=np.float16)
  full[:, :, 1, 1] = w16[:, :, 0, 0]
  return full


def _finish(arch, feat, n_up, r, sections):
  head_bytes = 32 + 24 * len(sections)
  off = -(-head_bytes // 256) * 256
  directory, payload = b'', b''
  for kind, index, data in sections:
    directory += struct.pack('<IIQQ', kind, index, off + len(payload), len(data))
    payload += data + b'\0' * (-len(data) % 256)
  header = struct.pack('<8I', MAGIC, VERSION, arch, feat, n_up, r, len(sections), 0)
  blob = header + directory
  blob += b'\0' * (off - len(blob)) + payload
  return arch, blob


def pack_lite(sd):
  """MoeNet_lite2.Net: 1x1 convolutions ride the 3x3 kernels as centre-tap filters; 48 channels zero-padded to 64;
  branch 0 = `uim` (on `out`, head convt_I1), branch 1 = `ures` (on the trunk, head convt_R1) — the same roles as
  `u` / `convt_R1` of MyNet; FRM gates (models.py:270-287) go to their own sections."""
  feat = 48
  n_up = len([k for k in sd if k.startswith('ures.') and k.endswith('.0.weight')])
  sections = []
  first = np.zeros((9, 64), dtype=np.float32)
  first[4, :feat] = _h(sd['conv_input.weight']).astype(np.float32).reshape(feat)
  sections.append((SEC_FIRST_W, 0, first.tobytes()))
  scalars = np.zeros(32, dtype=np.float32)
  scalars[0] = _h(sd['relu.weight']).astype(np.float32).reshape(-1)[0]
  sections.append((SEC_TRUNK_IMG, 0, conv_image(_center(_h(sd['conv_input2.weight']))).tobytes()))
  for b, name in enumerate(('convt_F11', 'convt_F12', 'convt_F13')):
    sections.append((SEC_TRUNK_IMG, 1 + 2 * b, conv_image(_h(sd[name + '.conv_1.weight'])).tobytes()))
    sections.append((SEC_TRUNK_IMG, 2 + 2 * b, conv_image(_h(sd[name + '.conv_2.weight'])).tobytes()))
    scalars[1 + 1 + 2 * b] = _h(sd[name + '.relu.weight']).astype(np.float32).reshape(-1)[0]
    frm = np.zeros(3 * 64 + 4 + 64 * 4 + 64, dtype=np.float32)      # w0[3][64], b0[4], w1[64][4], b1[64]
    frm[:192].reshape(3, 64)[:, :feat] = _h(sd[name + '.se.conv_du.0.weight']).astype(np.float32).reshape(3, feat)
    frm[192:195] = _h(sd[name + '.se.conv_du.0.bias']).astype(np.float32)
    frm[196:452].reshape(64, 4)[:feat, :3] = _h(sd[name + '.se.conv_du.2.weight']).astype(np.float32).reshape(feat, 3)
    frm[452:452 + feat] = _h(sd[name + '.se.conv_du.2.bias']).astype(np.float32)
    sections.append((SEC_FRM, b, frm.tobytes()))
  for bi, (name, head) in enumerate((('uim', 'convt_I1'), ('ures', 'convt_R1'))):
    for st in range(n_up):
      w = _center(_h(sd['%s.%d.0.weight' % (name, st)]))            # (192, 48, 3, 3)
      bias = _h(sd['%s.%d.0.bias' % (name, st)]).astype(np.float32)
      imgs, bs = [], []
      for i in range(2):
        for j in range(2):
          sel = np.arange(feat) * 4 + i * 2 + j
          imgs.append(conv_image(w[sel]))
          bq = np.zeros(64, dtype=np.float32)
          bq[:feat] = bias[sel]
          bs.append(bq)
      sections.append((SEC_UP_IMG, 4 * bi + st, np.concatenate(imgs).tobytes()))
      sections.append((SEC_UP_BIAS, 4 * bi + st, np.stack(bs).tobytes()))
      scalars[14 + 4 * bi + st] = _h(sd['%s.%d.2.weight' % (name, st)]).astype(np.float32).reshape(-1)[0]
    hw = np.zeros((9, 64), dtype=np.float32)
    hw[4, :feat] = _h(sd[head + '.weight']).astype(np.float32).reshape(feat)
    sections.append((SEC_HEAD_W, bi, hw.tobytes()))
  sections.append((SEC_SCALARS, 0, scalars.tobytes()))
  return _finish(ARCH_LITE, feat, n_up, 2, sections)


def pack(sd):
  """sd: checkpoint state dict (torch tensors or numpy arrays).  Returns (arch, bytes)."""
  arch = detect_arch(sd)
  if arch == ARCH_LITE:
    return pack_lite(sd)
  feat = int(sd['conv_input.weight'].shape[0])
  n_up, r = {ARCH_NETDN: (0, 0), ARCH_NET2X: (1, 2), ARCH_NET3X: (1, 3), ARCH_NET4X: (2, 2)}[arch]
  sections = []   # (kind, index, bytes)

  first = np.zeros((9, 64), dtype=np.float32)
  first[:, :feat] = _h(sd['conv_input.weight']).astype(np.float32).reshape(feat, 9).T
  sections.append((SEC_FIRST_W, 0, first.tobytes()))

  scalars = np.zeros(32, dtype=np.float32)
  scalars[0] = _h(sd['relu.weight']).astype(np.float32).reshape(-1)[0]
  sections.append((SEC_TRUNK_IMG, 0, conv_image(_h(sd['conv_input2.weight'])).tobytes()))
  for b in range(6):
    p = 'convt_F%d.0.' % (b + 1)
    sections.append((SEC_TRUNK_IMG, 1 + 2 * b, conv_image(_h(sd[p + 'conv_1.weight'])).tobytes()))
    sections.append((SEC_TRUNK_IMG, 2 + 2 * b, conv_image(_h(sd[p + 'conv_2.weight'])).tobytes()))
    scalars[1 + 1 + 2 * b] = _h(sd[p + 'relu.weight']).astype(np.float32).reshape(-1)[0]
    scalars[1 + 2 + 2 * b] = _h(sd[p + 'scale.scale']).astype(np.float32).reshape(-1)[0]

  for bi, name in enumerate(('u', 'convt_R1')):
    for s in range(n_up):
      w = _h(sd['%s.%d.0.weight' % (name, s)])               # (64*r*r, 64, 3, 3)
      bias = _h(sd['%s.%d.0.bias' % (name, s)]).astype(np.float32)
      imgs, bs = [], []
      for i in range(r):
        for j in range(r):
          sel = np.arange(64) * r * r + i * r + j            # PixelShuffle: channel c*r*r+i*r+j -> (c, i, j)
          imgs.append(conv_image(w[sel]))
          bs.append(bias[sel])
      sections.append((SEC_UP_IMG, 4 * bi + s, np.concatenate(imgs).tobytes()))
      sections.append((SEC_UP_BIAS, 4 * bi + s, np.stack(bs).astype(np.float32).tobytes()))
      scalars[14 + 4 * bi + s] = _h(sd['%s.%d.2.weight' % (name, s)]).astype(np.float32).reshape(-1)[0]
    hk = ('%s.%d.weight' % (name, n_up)) if n_up else (name + '.weight')
    head = np.zeros((9, 64), dtype=np.float32)
    head[:, :feat] = _h(sd[hk]).astype(np.float32).reshape(feat, 9).T   # (1,F,3,3) -> [tap][cin]
    sections.append((SEC_HEAD_W, bi, head.tobytes()))
  sections.append((SEC_SCALARS, 0, scalars.tobytes()))

  return _finish(arch, feat, n_up, r, sections)
