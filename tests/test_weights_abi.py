"""Weight packing and the C-ABI surface, checked without a GPU."""
import ctypes
import os
import re
import struct

import numpy as np
import pytest
import torch

import helpers as H
from moephoto_b200 import _lib, weights as W


def _unswizzle(img):
  a = np.frombuffer(img, dtype=np.float16).reshape(9, 64, 8, 8)
  out = np.empty_like(a)
  for row in range(64):
    for g in range(8):
      out[:, row, g ^ (row & 7)] = a[:, row, g]
  return out.reshape(9, 64, 64)          # [tap][cout][cin]


def test_conv_image_is_the_swizzled_k_major_layout():
  rng = np.random.default_rng(0)
  w = rng.standard_normal((64, 64, 3, 3)).astype(np.float16)
  img = W.conv_image(w)
  assert img.nbytes == 73728
  back = _unswizzle(img.tobytes())
  assert np.array_equal(back, w.reshape(64, 64, 9).transpose(2, 0, 1))
  w48 = rng.standard_normal((48, 48, 3, 3)).astype(np.float16)
  back = _unswizzle(W.conv_image(w48).tobytes())
  assert np.array_equal(back[:, :48, :48], w48.reshape(48, 48, 9).transpose(2, 0, 1))
  assert not back[:, 48:].any() and not back[:, :, 48:].any()


@pytest.mark.parametrize('key,arch,n_up,r', [('a2', 2, 1, 2), ('a3', 3, 1, 3), ('a4', 4, 2, 2), ('dn_lite15', 1, 0, 0), ('lite2', 5, 1, 2),
                                             ('lite8', 5, 3, 2)])
def test_blob_layout(key, arch, n_up, r):
  sd = H.load_weights(key)
  a, blob = W.pack(sd)
  assert a == arch
  magic, ver, arch_, feat, n_up_, r_, nsec, _ = struct.unpack_from('<8I', blob, 0)
  assert (magic, ver, arch_, n_up_, r_) == (0x42454F4D, 2, arch, n_up, r) and feat == (48 if arch in (1, 5) else 64)
  kinds = {}
  for i in range(nsec):
    kind, index, off, nb = struct.unpack_from('<IIQQ', blob, 32 + 24 * i)
    assert off % 256 == 0 and off + nb <= len(blob)
    kinds.setdefault(kind, {})[index] = (off, nb)
  assert len(kinds[W.SEC_TRUNK_IMG]) == (7 if arch == 5 else 13) and len(kinds[W.SEC_HEAD_W]) == 2
  if arch == 5:
    assert len(kinds[W.SEC_FRM]) == 3 and set(kinds[W.SEC_UP_IMG]) == {4 * b + st for b in range(2) for st in range(n_up)}
    return
  assert len(kinds.get(W.SEC_UP_IMG, {})) == 2 * n_up
  # PixelShuffle permutation: image (i,j) of the first upsample conv holds channels c*r*r + i*r + j
  if n_up:
    off, nb = kinds[W.SEC_UP_IMG][0]
    assert nb == r * r * 73728
    q = r + 1 if r == 3 else 3           # some sub-pixel (i,j) != (0,0)
    img = _unswizzle(blob[off + q * 73728: off + (q + 1) * 73728])
    w = sd['u.0.0.weight'].astype(np.float16)
    sel = np.arange(64) * r * r + q
    assert np.array_equal(img, w[sel].reshape(64, 64, 9).transpose(2, 0, 1))
    boff, _ = kinds[W.SEC_UP_BIAS][0]
    bias = np.frombuffer(blob, dtype=np.float32, count=r * r * 64, offset=boff).reshape(r * r, 64)
    assert np.array_equal(bias[q], sd['u.0.0.bias'].astype(np.float16).astype(np.float32)[sel])
  off, _ = kinds[W.SEC_SCALARS][0]
  sc = np.frombuffer(blob, dtype=np.float32, count=32, offset=off)
  assert sc[0] == np.float32(np.float16(sd['relu.weight'][0]))
  assert sc[3] == np.float32(np.float16(sd['convt_F1.0.scale.scale'][0]))


def test_pack_accepts_torch_state_dicts_and_rejects_others():
  sd = {k: torch.from_numpy(v) for k, v in H.load_weights('a2').items()}
  assert W.pack(sd)[1] == W.pack(H.load_weights('a2'))[1]
  bad = dict(sd)
  bad['conv_input.weight'] = torch.zeros(32, 1, 3, 3)
  with pytest.raises(ValueError):
    W.pack(bad)


def test_library_exports_every_declared_symbol():
  hdr = open(os.path.join(H.ROOT, 'include', 'moephoto_b200.h')).read()
  hdr = re.sub(r'/\*.*?\*/', '', hdr, flags=re.S)
  declared = set(re.findall(r'\b(moe_[a-z0-9_]+)\s*\(', hdr))
  assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
  lib = _lib.load()
  for name in declared:
    assert getattr(lib, name) is not None
  assert lib.moe_abi_version() == 1


def test_no_cpu_fallback():
  """without a GPU the engine refuses to exist, and doCrop refuses host tensors"""
  lib = _lib.load()
  if not torch.cuda.is_available():
    h = ctypes.c_void_p()
    assert lib.moe_engine_create(0, ctypes.byref(h)) == _lib.MOE_ERR_NO_DEVICE
    assert b'no CPU fallback' in lib.moe_last_error()
    with pytest.raises(RuntimeError):
      _lib.check(_lib.MOE_ERR_NO_DEVICE)
  from moephoto_b200 import imageProcess as IP
  with pytest.raises(RuntimeError):
    IP.run_plan(None, torch.zeros(3, 8, 8), IP.TilePlan.single(8, 8, 2))


def test_product_never_imports_the_oracle():
  pkg = os.path.join(H.ROOT, 'moephoto_b200')
  for dirpath, _, files in os.walk(pkg):
    for f in files:
      if f.endswith(('.py', '.cu', '.cuh', '.h')):
        src = open(os.path.join(dirpath, f)).read()
        assert not re.search(r'^\s*(from|import)\s+oracle\b', src, flags=re.M), f
        assert 'refharness' not in src, f
