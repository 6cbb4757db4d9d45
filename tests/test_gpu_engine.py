"""Parity tests proper (-m gpu): the CUDA path, called through the reference-shaped API and the C ABI,
against the oracle and the golden vectors of the unmodified reference.

Tolerances, all stated here:
  * vs the reference's OWN fp16 output (goldens `<case>.ref16`: the unmodified reference in its GPU fp16 configuration,
    executed on CPU) and vs the oracle in mode 'ref16' (same rounding points, restated in numpy/C) — helpers.assert_ref16_bar:
      smooth images        max-abs <= 1e-3 (one fp16 ulp at 1.0 is 9.77e-4) and PSNR >= 75 dB — BASELINE's "1e-3 max-abs fp16";
      uniform white noise  max-abs <= 2e-3 (two ulps), <= 0.2 % of the pixels beyond 1e-3, PSNR >= 68 dB.
    The white-noise bar is not a loosening for the engine: it is how far two exact implementations of the SAME rounding
    contract are apart when only the fp32 summation order inside conv2d differs (oracle vs reference, CPU only:
    profiles/r02_parity_oracle_vs_reference.txt; the reference's cuDNN path vs its CPU path: test_reference_gpu_port_*).
  * vs the reference's fp32 CPU output (goldens `.ref`): PSNR >= 60 dB (the north-star bar).
  * integer frame conversions, band sharding, determinism: bit-exact.
"""
import ctypes
import os

import numpy as np
import pytest
import torch

import helpers as H

pytestmark = pytest.mark.gpu


def _conv_ref(x, w, bias, r, epi, param, skip):
  """fp32 convolution, then the epilogue with the reference's per-op fp16 roundings (csrc/conv_tc.cuh::epi_apply)"""
  q = lambda t: t.half().float()
  xn = x.float().permute(0, 3, 1, 2)
  y = torch.nn.functional.conv2d(xn, w.float(), None if bias is None else bias.float(), padding=1)
  if epi == 1:
    y = q(y)
    y = torch.where(y >= 0, y, param * y)
  elif epi == 2:
    y = skip.float().permute(0, 3, 1, 2) + q(param * q(y))
  elif epi == 3:
    y = q(torch.nn.functional.pixel_shuffle(y, r) if r > 1 else y)
    y = torch.where(y >= 0, y, param * y)
  return y.permute(0, 2, 3, 1).contiguous()


def _conv_ref_biased(x, w, bias, r, param):
  """EPI_BIAS_PRELU with the GPU's two ops: q(q(conv) + bias), PixelShuffle, PReLU"""
  q = lambda t: t.half().float()
  y = torch.nn.functional.conv2d(x.float().permute(0, 3, 1, 2), w.float(), None, padding=1)
  y = q(q(y) + bias.float().view(1, -1, 1, 1))
  y = torch.nn.functional.pixel_shuffle(y, r) if r > 1 else y
  y = torch.where(y >= 0, y, param * y)
  return y.permute(0, 2, 3, 1).contiguous()


def _run_conv(eng, x, w, bias, r, epi, param, skip):
  from moephoto_b200 import _lib, weights as W
  n, h, wd, _ = x.shape
  w16 = w.half().cpu().numpy()
  imgs, bs = [], []
  for i in range(r):
    for j in range(r):
      sel = np.arange(64) * r * r + i * r + j
      imgs.append(W.conv_image(w16[sel]))
      if bias is not None:
        bs.append(bias.float().cpu().numpy()[sel])
  img = torch.from_numpy(np.concatenate(imgs)).cuda()
  bdev = torch.from_numpy(np.stack(bs).astype(np.float32)).cuda() if bias is not None else None
  out = skip.clone() if epi == 2 else torch.full((n, h * r, wd * r, 64), float('nan'), dtype=torch.half, device='cuda')
  _lib.check(eng.lib.moe_conv3x3_c64(eng.handle, ctypes.c_void_p(x.data_ptr()), ctypes.c_void_p(out.data_ptr()),
                                     ctypes.c_void_p(out.data_ptr()) if epi == 2 else None, ctypes.c_void_p(img.data_ptr()),
                                     ctypes.c_void_p(bdev.data_ptr()) if bdev is not None else None, n, h, wd, r, epi, float(param), None))
  torch.cuda.synchronize()
  return out


CONV_CASES = [(1, 1, 1, 1, 0), (1, 8, 128, 1, 0), (1, 20, 129, 1, 1), (2, 33, 200, 1, 2), (3, 40, 48, 1, 1), (1, 17, 300, 2, 3),
              (1, 12, 130, 3, 3), (3, 70, 257, 1, 0), (1, 3, 1000, 1, 2), (2, 300, 127, 1, 1), (1, 9, 7, 2, 3)]


@pytest.mark.parametrize('n,h,w,r,epi', CONV_CASES)
def test_conv3x3_tcgen05_matches_fp32_reference_and_simt(engine, n, h, w, r, epi):
  """one layer: empty-ish, ragged (w not a multiple of the 128-px strip), every epilogue, PixelShuffle 2 and 3"""
  g = torch.Generator().manual_seed(n * 7919 + h * 31 + w)
  x = (torch.randn(n, h, w, 64, generator=g) * 0.5).half().cuda()
  wt = (torch.randn(64 * r * r, 64, 3, 3, generator=g) * 0.05).half()
  bias = (torch.randn(64 * r * r, generator=g) * 0.1).half() if epi == 3 else None
  skip = torch.randn(n, h, w, 64, generator=g).half().cuda() if epi == 2 else None
  ref = _conv_ref_biased(x, wt.cuda(), bias.cuda(), r, 0.25) if epi == 3 else _conv_ref(x, wt.cuda(), None, r, epi, 0.25, skip)
  ref = ref.half().float()                                            # what a correct implementation stores
  engine.set_conv_path(simt=False)
  tc = _run_conv(engine, x, wt, bias, r, epi, 0.25, skip).float()
  engine.set_conv_path(simt=True)
  simt = _run_conv(engine, x, wt, bias, r, epi, 0.25, skip).float()
  engine.set_conv_path(simt=False)
  assert not torch.isnan(tc).any() and not torch.isnan(simt).any()
  # a flipped intermediate rounding (fp32 summation order) moves the stored value by one fp16 ulp of the result, and that ulp
  # through the epilogue's second op (bias add / x scale + skip / PReLU) by at most one more
  tol = 2.0 ** -9 * torch.clamp(ref.abs(), min=1.0) + 1e-4
  assert ((tc - ref).abs() <= tol).all()
  assert ((simt - ref).abs() <= tol).all()
  # same rounding points, different fp32 summation order: equal up to one fp16 ulp on a small fraction
  d = (tc - simt).abs()
  assert (d <= 2.0 ** -9 * torch.clamp(ref.abs(), min=1.0)).all(), (d / torch.clamp(ref.abs(), min=1.0)).max().item()
  assert (d > 0).float().mean() < 0.05


@pytest.mark.parametrize('name', H.case_names())
def test_golden_cases_through_the_reference_api(engine, name):
  """every golden through runSR.getOpt / runSR.sr / runDN.getOpt / RGBFilter.
  (1) DEFAULT numerics (the reference's GPU path: a biased convolution is two ops) against the oracle in mode 'ref16' — the
      tight bar — and against the CPU-executed golden — the cross-platform bar (helpers.assert_cross_platform_bar);
  (2) with the engine's biased convolutions switched to the CPU execution's single rounding (bias_fused), against the golden
      the UNMODIFIED reference produced in exactly that arithmetic — the tight bar, no oracle in between."""
  c = H.load_case(name)
  y = H.run_case_engine(c)                  # asserts the tile plan equals the reference's too
  assert y.shape == c['ref'].shape
  orc = H.assert_ref16_bar(y, c, ref=H.run_case_oracle(c, mode='ref16'), what='engine vs oracle ref16')
  xp = H.assert_cross_platform_bar(y, c, what='engine (GPU bias semantics) vs the CPU-executed reference fp16 golden')
  yc = H.run_case_engine(c, cpu_bias=True)
  got = H.assert_ref16_bar(yc, c, what='engine (bias_fused) vs the reference fp16 golden')
  print('PARITY %-16s bias_fused vs reference fp16 golden: max %.2e mean %.2e PSNR %.1f dB | default vs oracle ref16: max %.2e PSNR %.1f dB | '
        'default vs golden: max %.2e PSNR %.1f dB | vs reference fp32: PSNR %.1f dB'
        % (name, got[0], got[1], got[2], orc[0], orc[2], xp[0], xp[2], H.psnr(y, c['ref'])))
  # the north-star bar; on uniform white noise the reference's OWN fp16 output is only 60.1 dB (lite4) - 69 dB from its fp32 output
  assert H.psnr(y, c['ref']) >= (59.0 if H.is_white_noise(c) else 60.0)
  if c['alpha'] is not None:
    assert np.array_equal(y[3], c['alpha'].astype(np.float16).astype(np.float32))


@pytest.mark.parametrize('name', ['a2_tiled', 'a4_tiled', 'dn15_tiled', 'lite4_single'])
def test_simt_cross_check_path_agrees_with_tensor_core_path(engine, name):
  c = H.load_case(name)
  y_tc = H.run_case_engine(c)
  engine.set_conv_path(simt=True)
  try:
    y_simt = H.run_case_engine(c)
  finally:
    engine.set_conv_path(simt=False)
  assert np.abs(y_tc - y_simt).max() <= 1e-3


def _sr_opt(key, scale, crop=0, ram=int(178 * 2 ** 30 * .9)):
  from moephoto_b200 import runSR
  from moephoto_b200.config import config
  config.freeMemOverride, config.crop_sr = ram, (crop or 'auto')
  try:
    return runSR.getOpt({'model': 'lite' if key.startswith('lite') else 'a', 'scale': scale}, weights=H.load_weights(key))
  finally:
    config.crop_sr = 'auto'


@pytest.mark.parametrize('key,scale,shape,crop', [('a2', 2, (3, 300, 420), 128), ('a4', 4, (3, 200, 333), 96), ('a3', 3, (2, 150, 260), 0),
                                                  ('lite4', 4, (3, 120, 200), 64)])
def test_row_band_sharding_is_bit_exact(engine, key, scale, shape, crop):
  """moe_run_plan's row window (what each GPU of a multi-GPU run computes) reproduces the full run exactly"""
  from moephoto_b200 import imageProcess as IP, parallel as PAR
  from moephoto_b200.config import config
  opt = _sr_opt(key, scale, crop)
  try:
    x = torch.rand(shape, generator=torch.Generator().manual_seed(5)).half().cuda()
    full = IP.doCrop(opt, x)
    for world in (2, 3, 8):
      out = torch.full_like(full, float('nan'))
      for rank in range(world):
        lo, hi = PAR.band_rows(shape[1], scale, world, rank)
        IP.run_plan(opt.modelCached, x, opt.plan, out, rows=(lo, hi))
      assert torch.equal(out, full), world
  finally:
    config.freeMemOverride = None


def test_full_size_properties_1080p_a2(engine):
  """BASELINE configs[1] (1920x1080 -> 3840x2160, a2): determinism, sharding invariance on the real frame size,
  the reference's single-tile auto plan, and tile-count independence within the seam tolerance"""
  from moephoto_b200 import imageProcess as IP, parallel as PAR
  from moephoto_b200.config import config
  opt = _sr_opt('a2', 2)
  try:
    g = torch.Generator().manual_seed(11)
    lo_res = torch.rand(3, 135, 240, generator=g)
    x = torch.nn.functional.interpolate(lo_res[None], size=(1080, 1920), mode='bicubic')[0].clamp(0, 1).half().cuda()
    y1 = IP.doCrop(opt, x)
    assert len(opt.plan.tiles) == 1 and tuple(y1.shape) == (3, 2160, 3840)
    assert torch.equal(y1, IP.doCrop(opt, x))
    assert torch.isfinite(y1).all() and float(y1.min()) > -0.2 and float(y1.max()) < 1.2
    out = torch.empty_like(y1)
    for rank in range(4):
      lo, hi = PAR.band_rows(1080, 2, 4, rank)
      IP.run_plan(opt.modelCached, x, opt.plan, out, rows=(lo, hi))
    assert torch.equal(out, y1)
    # a 512-px crop gives 12 tiles; away from tile borders the receptive field (16 px) sees the same data
    opt12 = _sr_opt('a2', 2, crop=512)
    y12 = IP.doCrop(opt12, x)
    assert len(opt12.plan.tiles) == 12
    assert torch.equal(y12[:, 100:800, 100:800], y1[:, 100:800, 100:800])
    assert 10 * np.log10(1.0 / float(((y12.float() - y1.float()) ** 2).mean())) > 60.0
  finally:
    config.freeMemOverride = None


def test_bare_network_call_matches_oracle(engine):
  """opt(x) on a (N,1,h,w) tile, like the reference's Option.__call__ (imageProcess.py:391-395)"""
  from oracle import net as N
  opt = _sr_opt('a4', 4)
  from moephoto_b200.config import config
  config.freeMemOverride = None
  sd = H.load_weights('a4')
  x = torch.rand(2, 1, 37, 53, generator=torch.Generator().manual_seed(2)).half()
  y = opt(x.cuda())
  assert tuple(y.shape) == (2, 1, 148, 212)
  want = N.forward(sd, x.float().numpy(), mode='ref16')
  d = np.abs(y.float().cpu().numpy() - want)
  assert d.max() <= 2e-3 and (d > 1e-3).mean() <= 2e-3          # the white-noise bar (module docstring)


def test_frame_conversions_are_bit_exact(engine):
  from oracle import tiling as T
  from moephoto_b200 import imageProcess as IP
  rng = np.random.default_rng(0)
  img8 = rng.integers(0, 256, (37, 51, 3), dtype=np.uint8)
  img16 = rng.integers(0, 65536, (37, 51, 3), dtype=np.uint16)
  a = IP.toTorch(8)(img8)
  assert np.array_equal(a.cpu().numpy(), T.to_planar(img8, 8, np.float16))
  b = IP.toTorch(16)(img16)
  assert np.array_equal(b.cpu().numpy(), T.to_planar(img16, 16, np.float16))
  assert np.array_equal(IP.toTorch(8, swapRB=True)(img8).cpu().numpy(), T.to_planar(img8[:, :, ::-1], 8, np.float16))
  y = torch.from_numpy(rng.uniform(-.1, 1.1, (3, 37, 51)).astype(np.float16)).cuda()
  for bits in (8, 16):
    got = IP.toOutput(bits)(y)
    want = T.to_output(y.cpu().numpy(), bits)
    assert np.array_equal(got.astype(np.int64), want.astype(np.int64) & (0xFFFF if bits == 16 else 0xFF))
  assert np.array_equal(IP.toOutput(8)(IP.toFloat(y)), IP.toOutput(8)(y))      # accepts what toFloat hands over


def test_strength_and_alpha(engine):
  """RGBFilter with strength != 1 (strengthOp, imageProcess.py:562), fp16 rounding per op"""
  from moephoto_b200 import runDN, imageProcess as IP
  from moephoto_b200.config import config
  config.freeMemOverride = int(4e9)
  try:
    opt = runDN.getOpt({'model': 'lite15', 'strength': 0.6}, weights=H.load_weights('dn_lite15'))
    x = torch.rand(4, 50, 66, generator=torch.Generator().manual_seed(3)).half().cuda()
    y = IP.RGBFilter(opt)(x)
    opt1 = runDN.getOpt({'model': 'lite15'}, weights=H.load_weights('dn_lite15'))
    y1 = IP.RGBFilter(opt1)(x)
    h16 = lambda t: t.half().float()
    want = h16(h16(0.6 * y1[:3].float()) + h16(np.float32(1 - 0.6) * x[:3].float()))
    assert torch.equal(y[:3].float(), want) and torch.equal(y[3], x[3])
  finally:
    config.freeMemOverride = None


def test_ensemble_is_the_average_of_the_dihedral_passes(engine):
  """structure check; the values are checked against the reference by the goldens a2_ens3 / a2_ens7"""
  from moephoto_b200 import runSR, imageProcess as IP
  from moephoto_b200.config import config
  config.freeMemOverride = int(4e9)
  try:
    sd = H.load_weights('a2')
    x = torch.rand(3, 48, 72, generator=torch.Generator().manual_seed(4)).half().cuda()
    o0 = runSR.getOpt({'model': 'a', 'scale': 2, 'ensemble': 0}, weights=sd)
    o3 = runSR.getOpt({'model': 'a', 'scale': 2, 'ensemble': 3}, weights=sd)
    y3 = runSR.sr(o3)(x)
    f = lambda t: IP.doCrop(o0, t.contiguous())
    ot = runSR.getOpt({'model': 'a', 'scale': 2}, weights=sd)
    ft = lambda t: IP.doCrop(ot, t.contiguous())
    want = f(x) + ft(x.transpose(-1, -2)).transpose(-1, -2) + f(x.flip(-1)).flip(-1) + f(x.flip(-1, -2)).flip(-1, -2)
    assert torch.allclose(y3.float(), (want / 4).float(), atol=2e-3)
  finally:
    config.freeMemOverride = None


def test_host_buffer_entry_point_equals_the_api_path(engine):
  """moe_enhance_host (uint8 host frame in, uint8 host frame out) == toTorch -> doCrop -> toOutput"""
  from moephoto_b200 import _lib, imageProcess as IP
  from moephoto_b200.config import config
  opt = _sr_opt('a2', 2, crop=64, ram=int(4e9))
  try:
    img = np.random.default_rng(9).integers(0, 256, (90, 130, 3), dtype=np.uint8)
    want = IP.toOutput(8)(IP.doCrop(opt, IP.toTorch(8)(img)))
    out = np.empty((180, 260, 3), dtype=np.uint8)
    _lib.check(engine.lib.moe_enhance_host(opt.modelCached.handle, img.ctypes.data_as(ctypes.c_void_p), 8, ctypes.byref(opt.plan.c),
                                           out.ctypes.data_as(ctypes.c_void_p), 8, None))
    assert np.array_equal(out, want)
  finally:
    config.freeMemOverride = None


def test_host_buffer_entry_point_overlapped_column_copies(engine):
    """one row of column strips (the shape of the reference's plan for big frames): moe_enhance_host converts and copies
    finished columns on a second stream while the next tile computes — same bytes as the plain path, 8- and 16-bit"""
    from moephoto_b200 import _lib, imageProcess as IP
    from moephoto_b200.config import config
    opt = _sr_opt('a4', 4, crop=64, ram=int(4e9))
    try:
        rng = np.random.default_rng(12)
        img = rng.integers(0, 256, (48, 300, 3), dtype=np.uint8)
        y = IP.doCrop(opt, IP.toTorch(8)(img))
        assert len(opt.plan.tiles) >= 4 and all(t[0] == 0 for t in opt.plan.tiles)          # a single tile row
        for bits, dt in ((8, np.uint8), (16, np.uint16)):
            want = IP.toOutput(bits)(y)
            out = np.zeros((192, 1200, 3), dtype=dt)
            _lib.check(engine.lib.moe_enhance_host(opt.modelCached.handle, img.ctypes.data_as(ctypes.c_void_p), 8, ctypes.byref(opt.plan.c),
                                                   out.ctypes.data_as(ctypes.c_void_p), bits, None))
            assert np.array_equal(out, want)
    finally:
        config.freeMemOverride = None


def test_errors_are_status_codes_not_crashes(engine):
  from moephoto_b200 import _lib, imageProcess as IP
  opt = _sr_opt('a2', 2, ram=int(4e9))
  from moephoto_b200.config import config
  config.freeMemOverride = None
  x = torch.rand(3, 40, 40).half().cuda()
  with pytest.raises(ValueError):
    IP.run_plan(opt.modelCached, x, IP.TilePlan.single(40, 40, 4))          # scale mismatch
  with pytest.raises(ValueError):
    IP.run_plan(opt.modelCached, x, IP.TilePlan.single(48, 40, 2))          # shape mismatch
  bad = IP.TilePlan([(0, 400, 0, 40, 0, 0, 800, 80)], 2, 0, 40, 40, 0, 0, np.zeros(1, np.float32))
  with pytest.raises(ValueError):
    IP.run_plan(opt.modelCached, x, bad)                                     # tile outside the image
  h = ctypes.c_void_p()
  assert engine.lib.moe_model_load(engine.handle, 2, b'garbage-garbage-garbage-garbage-garbage', 40, ctypes.byref(h)) == _lib.MOE_ERR_INVALID
  with pytest.raises(MemoryError):
    cfgopt = _sr_opt('a2', 2, ram=1000)
    IP.doCrop(cfgopt, x)
  config.freeMemOverride = None
  # crafted blobs: a header that does not describe its architecture, a section whose offset + size wraps around
  import struct
  from moephoto_b200 import weights as W
  arch, blob = W.pack(H.load_weights('a2'))
  load = lambda bts: engine.lib.moe_model_load(engine.handle, arch, (ctypes.c_uint8 * len(bts)).from_buffer_copy(bts), len(bts), ctypes.byref(h))
  lied = bytearray(blob)
  struct.pack_into('<I', lied, 16, 3)                                        # n_up = 3 in a Net2x blob
  assert load(bytes(lied)) == _lib.MOE_ERR_INVALID
  wrap = bytearray(blob)
  struct.pack_into('<Q', wrap, 32 + 8, 2 ** 64 - 256)                        # first section: offset just below 2^64
  assert load(bytes(wrap)) == _lib.MOE_ERR_INVALID and b'bad blob section' in engine.lib.moe_last_error()
  assert load(blob) == _lib.MOE_OK                                           # the engine is unharmed
  engine.lib.moe_model_free(h)


def test_chained_dn_then_sr_on_a_frame_batch_16bit_route(engine):
  """BASELINE configs[3] in miniature: two 16-bit frames (bgr48le as video.py pipes them) -> toTorch(16) ->
  DN lite15 -> SR a2 -> toOutput(16), frames batched as planes (bit-identical to per-frame calls
  when cropsize is pinned), checked against the oracle (mode 'ref16')."""
  from oracle import net as N, tiling as T
  from moephoto_b200 import runSR, runDN, imageProcess as IP
  from moephoto_b200.config import config
  rng = np.random.default_rng(21)
  base = (np.clip(np.add.outer(np.linspace(0.1, 0.8, 40), np.linspace(0.0, 0.2, 56)), 0, 1) * 65535)
  frames = [np.clip(base[:, :, None] + rng.normal(0, 900, (40, 56, 3)), 0, 65535).astype(np.uint16) for _ in range(2)]
  config.freeMemOverride, config.crop_dn, config.crop_sr = int(4e9), 32, 32
  try:
    odn = runDN.getOpt({'model': 'lite15'}, weights=H.load_weights('dn_lite15'))
    osr = runSR.getOpt({'model': 'a', 'scale': 2}, weights=H.load_weights('a2'))
    x = torch.cat([IP.toTorch(16, swapRB=True)(f) for f in frames], 0)           # (6,40,56): BGR->RGB on load
    y = runSR.sr(osr)(IP.RGBFilter(odn)(x))
    assert tuple(y.shape) == (6, 80, 112)
    out = [IP.toOutput(16, swapRB=True)(y[3 * i:3 * i + 3]) for i in range(2)]
    # per-frame calls give the same bits as the batch
    y0 = runSR.sr(runSR.getOpt({'model': 'a', 'scale': 2}, weights=H.load_weights('a2')))(
        IP.RGBFilter(runDN.getOpt({'model': 'lite15'}, weights=H.load_weights('dn_lite15')))(x[:3]))
    assert torch.equal(y0, y[:3])
    # oracle chain
    sdn, ssr = H.load_weights('dn_lite15'), H.load_weights('a2')
    for i, f in enumerate(frames):
      xi = T.to_planar(f[:, :, ::-1], 16, np.float16).astype(np.float32)
      pd = T.make_plan((3, 40, 56), 4e9, .95 / 1253.4, 7, 1, 8, 32)
      d = T.rgb_filter(lambda a: N.forward(sdn, a, mode='ref16'), xi, pd, 1.0, np.float16).astype(np.float32)
      ps = T.make_plan((3, 40, 56), 4e9, .9 / 2473., 5, 2, 8, 32)
      s = T.do_crop(lambda a: N.forward(ssr, a, mode='ref16'), d, ps, np.float16)
      dd = np.abs(y[3 * i:3 * i + 3].float().cpu().numpy() - s.astype(np.float32))
      assert dd.max() <= 2e-3 and (dd > 1e-3).mean() <= 2e-3      # noisy 16-bit frames through two networks: the white-noise bar
      want = T.to_output(s.astype(np.float32), 16)[:, :, ::-1]
      assert np.abs(out[i].astype(np.int64) - (want.astype(np.int64) & 0xFFFF)).max() <= 140        # 2e-3 * 65536
  finally:
    config.freeMemOverride, config.crop_dn, config.crop_sr = None, 'auto', 'auto'


@pytest.mark.parametrize('flags', [dict(no_pair=True), dict(no_pair_trunk=True), dict(no_fuse=True), dict(static_sched=True), dict(no_arsb=True), dict(arsb_smem_mid=True), dict(arsb_solo=True)])
@pytest.mark.parametrize('name', ['a2_tiled', 'a3_tiled', 'a4_tiled', 'lite2_tiled', 'lite8_single'])
def test_every_tensor_core_kernel_variant_meets_the_same_bar(engine, name, flags):
    """the A/B switches keep the single-CTA conv kernel, the unfused CTA-pair kernel and head_tc_kernel alive;
    each variant is held to the same tolerance as the default path (CTA pairs + fused head)"""
    c = H.load_case(name)
    y_default = H.run_case_engine(c)
    engine.set_conv_path(**flags)
    try:
        y = H.run_case_engine(c)
    finally:
        engine.set_conv_path()
    H.assert_ref16_bar(y, c, ref=H.run_case_oracle(c, mode='ref16'), what='variant vs oracle ref16')
    assert H.psnr(y, c['ref']) >= 60.0
    assert np.abs(y - y_default).max() <= 1e-3
    if flags.get('static_sched') or flags.get('no_arsb') or flags.get('arsb_smem_mid') or flags.get('arsb_solo'):
        # who computes an item never changes its result; and the fused residual block has the rounding points AND the MMA
        # accumulation order of its two-launch form
        assert np.array_equal(y, y_default)


def test_fused_path_at_4k_tile_width(engine):
    """one reference tile of the bench workload shape in miniature height (3 x 64 x 968, a4): the CTA-pair kernels with
    an odd number of 128-px strips (968 = 7.56 strips -> 4 pairs, the last CTA half empty), checked against the SIMT path"""
    from moephoto_b200 import imageProcess as IP
    from moephoto_b200.config import config
    opt = _sr_opt('a4', 4)
    try:
        g = torch.Generator().manual_seed(17)
        x = torch.nn.functional.interpolate(torch.rand(1, 3, 8, 121, generator=g), size=(64, 968), mode='bicubic')[0].clamp(0, 1).half().cuda()
        y = IP.doCrop(opt, x)
        engine.set_conv_path(simt=True)
        try:
            y_simt = IP.doCrop(opt, x)
        finally:
            engine.set_conv_path()
        assert tuple(y.shape) == (3, 256, 3872)
        assert (y.float() - y_simt.float()).abs().max().item() <= 1e-3
    finally:
        config.freeMemOverride = None


@pytest.mark.parametrize('key,scale,shape', [('a2', 2, (3, 1, 1)), ('a2', 2, (1, 5, 7)), ('a4', 4, (3, 9, 3)), ('a3', 3, (2, 8, 130)), ('a2', 2, (4, 131, 9)),
                                             ('lite2', 2, (3, 3, 5)), ('lite8', 8, (1, 9, 131))])
def test_degenerate_and_ragged_shapes(engine, key, scale, shape):
    """smallest inputs the reference accepts (a single tile padded to 8 by reflect-then-zero padImage), widths just
    over one 128-px strip, plane counts 1..4 — against the oracle (mode 'ref16')"""
    from oracle import net as N, tiling as T
    from moephoto_b200 import imageProcess as IP
    from moephoto_b200.config import config
    opt = _sr_opt(key, scale, ram=int(4e9))
    try:
        x = torch.rand(shape, generator=torch.Generator().manual_seed(sum(shape))).half()
        y = IP.doCrop(opt, x.cuda())
        assert tuple(y.shape) == (shape[0], shape[1] * scale, shape[2] * scale)
        sd = H.load_weights(key)
        plan = T.make_plan(shape, int(4e9), opt.ramCoef, opt.padding, scale, 8, 0)
        assert plan.tiles == opt.plan.tiles and (plan.pad_h, plan.pad_w) == (opt.plan.pad_h, opt.plan.pad_w)
        want = T.do_crop(lambda a: N.forward(sd, a, mode='ref16'), x.float().numpy(), plan, np.float16).astype(np.float32)
        d = np.abs(y.float().cpu().numpy() - want)
        # the white-noise bar (module docstring); MoeNet_lite2's FRM gates: helpers.assert_ref16_bar
        assert d.max() <= (4e-3 if key.startswith('lite') else 2e-3) and (d > 1e-3).mean() <= (1e-2 if key.startswith('lite') else 2e-3)
    finally:
        config.freeMemOverride = None


def test_full_size_4k_a4_bench_workload(engine):
    """BASELINE configs[2] at full size (3840x2160 -> 15360x8640, a4, the reference's 4-strip auto plan):
      * 8-way row-band sharding (what 8 GPUs compute) is bit-identical to the single run;
      * a 96x96 crop taken well inside one reference tile and run on its own reproduces the full frame bit for bit in
        its interior (receptive field 16 LR px), and that crop matches the oracle — parity carried to the full size."""
    from oracle import net as N
    from moephoto_b200 import imageProcess as IP, parallel as PAR
    from moephoto_b200.config import config
    import bench
    opt = _sr_opt('a4', 4)
    try:
        x = IP.toTorch(8)(bench.synthetic_frame(2160, 3840, 0))
        y = IP.doCrop(opt, x)
        assert len(opt.plan.tiles) == 4 and tuple(y.shape) == (3, 8640, 15360)
        assert [(t[0], t[1], t[2], t[3]) for t in opt.plan.tiles] == [(0, 2160, 0, 968), (0, 2160, 963, 1931), (0, 2160, 1921, 2889), (0, 2160, 2880, 3840)]
        assert torch.isfinite(y).all()
        out = torch.empty_like(y)
        for rank in range(8):
            lo, hi = PAR.band_rows(2160, 4, 8, rank)
            IP.run_plan(opt.modelCached, x, opt.plan, out, rows=(lo, hi))
        assert torch.equal(out, y)
        del out
        # crop inside tile 1 (columns 963..1931), away from every seam
        cy, cx, s = 1000, 1400, 96
        crop = x[:, cy:cy + s, cx:cx + s].contiguous()
        yc = opt(crop.unsqueeze(1))[:, 0]                               # bare network on the crop, zero padded at its border
        m = 16
        assert torch.equal(yc[:, 4 * m:4 * (s - m), 4 * m:4 * (s - m)], y[:, 4 * (cy + m):4 * (cy + s - m), 4 * (cx + m):4 * (cx + s - m)])
        want = N.forward(H.load_weights('a4'), crop.float().cpu().numpy()[:, None], mode='ref16')[:, 0]
        assert np.abs(yc.float().cpu().numpy() - want).max() <= 1e-3
    finally:
        config.freeMemOverride = None


def test_frame_batched_video_route_equals_per_frame_calls(engine):
    """moephoto_b200.video.process_frames: B frames as 3B planes with the single-frame tile plan are bit-identical to the
    reference-style per-frame loop (video.py:351-360), for a DN -> SR chain on 16-bit BGR frames, on 1 and on 2 'ranks'"""
    from moephoto_b200 import runSR, runDN, imageProcess as IP, video
    from moephoto_b200.config import config
    rng = np.random.default_rng(5)
    frames = [rng.integers(0, 65536, (40, 56, 3), dtype=np.uint16) for _ in range(5)]
    config.freeMemOverride, config.crop_dn, config.crop_sr = int(4e9), 32, 32
    try:
        odn = runDN.getOpt({'model': 'lite15'}, weights=H.load_weights('dn_lite15'))
        osr = runSR.getOpt({'model': 'a', 'scale': 2}, weights=H.load_weights('a2'))
        want = []
        for f in frames:                                            # the reference's loop: one frame per iteration
            x = IP.toTorch(16, swapRB=True)(f)
            want.append(IP.toOutput(16, swapRB=True)(runSR.sr(osr)(IP.RGBFilter(odn)(x))))
        got = dict(video.process_frames(frames, [odn, osr], bit_depth=16, swap_rb=True, batch=3))
        assert sorted(got) == [0, 1, 2, 3, 4] and all(np.array_equal(got[i], want[i]) for i in range(5))
        halves = [dict(video.process_frames(frames, [odn, osr], batch=2, rank=r, world=2)) for r in range(2)]
        assert sorted(halves[0]) == [0, 2, 4] and sorted(halves[1]) == [1, 3]
        assert all(np.array_equal(halves[i % 2][i], want[i]) for i in range(5))
    finally:
        config.freeMemOverride, config.crop_dn, config.crop_sr = None, 'auto', 'auto'


@pytest.mark.parametrize('name', H.band_case_names())
def test_large_goldens_every_seam_band_against_the_reference(engine, name):
    """a4 on a 512x1024 frame cut into 3x6 tiles, and on a 96x3840 frame cut into the four column strips of the
    reference's 4K plan (same column anchors as the bench workload): every seam band (full length) and interior windows of
    the reference's fp16-configuration output, smooth-image bar"""
    c = H.load_band_case(name)
    y = H.run_case_engine(c, cpu_bias=True)          # the goldens are the CPU execution of the half model
    assert tuple(y.shape[1:]) == (c['img'].shape[0] * c['scale'], c['img'].shape[1] * c['scale'])
    worst, sq, npx = 0.0, 0.0, 0
    for (y0, y1, x0, x1), want in c['windows']:
        d = np.abs(y[:, y0:y1, x0:x1] - want)
        worst = max(worst, float(d.max()))
        sq += float((d.astype(np.float64) ** 2).sum())
        npx += d.size
    p = 10 * np.log10(1.0 / max(sq / npx, 1e-20))
    print('PARITY %-16s %d tiles, %d windows (%.1f MPix) vs reference fp16: max %.2e PSNR %.1f dB' % (name, len(c['tiles']), len(c['windows']), npx / 1e6, worst, p))
    assert worst <= 1e-3 and p >= 75.0


def test_blend_ramp_equals_the_reference_sigmoid_on_cuda_in_half(engine):
    """the reference evaluates the seam ramp on the GPU in half (imageProcess.py:109); the engine's host-side ramp
    (imageProcess.blendRamp) must give the same fp16 values for every seam width the served models produce"""
    from moephoto_b200 import imageProcess as IP
    for pad_sc in (7, 10, 14, 20, 27, 40):                        # dn 7x1, a2 5x2, (dn chained) 7x2, a4 / lite4 5x4, a3 9x3, lite8 5x8
        want = ((torch.arange(pad_sc, dtype=torch.half, device='cuda') / pad_sc - .5) * 9).sigmoid()
        got = torch.from_numpy(IP.blendRamp(pad_sc)).half()
        assert torch.equal(got, want.cpu()), pad_sc


@pytest.mark.parametrize('name', ['a2_tiled', 'a4_single', 'a3_tiled', 'dn15_tiled', 'lite4_single', 'a4_rand', 'lite4_rand'])
def test_reference_gpu_port_and_engine_are_equally_close_to_the_reference_goldens(engine, name):
    """/root/reference does not exist on the GPU box, so the reference's GPU arithmetic is reproduced by the oracle's
    PyTorch port run in half on CUDA (cuDNN conv2d, aten prelu / pixel_shuffle / add: op for op what the reference's
    nn.Modules launch; oracle/net.py::forward_torch) under the oracle's doCrop.  Three-way comparison against the golden
    the unmodified reference produced on CPU in the same fp16 configuration: the engine (default numerics = the GPU's) must
    meet the tight bar against the cuDNN port; how far port and engine are from the CPU-executed golden is printed."""
    from oracle import net as N, tiling as T
    c = H.load_case(name)
    y = H.run_case_engine(c)
    sd = H.load_weights(c['weights'])
    x = H.case_input(c, np.float16).astype(np.float32)
    net = lambda a: N.forward_torch(sd, a, dtype='float16', device='cuda')
    plan = H.oracle_plan(c)
    port = (T.do_crop(net, x, plan, np.float16) if c['kind'] == 'sr' else T.rgb_filter(net, x, plan, 1.0, np.float16)).astype(np.float32)
    de, dp, dep = np.abs(y - c['ref16']), np.abs(port - c['ref16']), np.abs(y - port)
    print('PARITY %-16s engine vs cuDNN-half port: max %.2e mean %.2e PSNR %.1f dB | port vs CPU golden: max %.2e mean %.2e PSNR %.1f dB | engine vs CPU golden: max %.2e mean %.2e PSNR %.1f dB'
          % (name, dep.max(), dep.mean(), H.psnr(y, port), dp.max(), dp.mean(), H.psnr(port, c['ref16']), de.max(), de.mean(), H.psnr(y, c['ref16'])))
    if not H.has_frm(c):
        H.assert_ref16_bar(y, c, ref=port, what='engine vs the reference GPU arithmetic (cuDNN half port)')
    else:
        # MoeNet_lite2: one flipped fp16 rounding of an FRM gate rescales a whole channel of a whole tile (models.py:287), so
        # two exact implementations are further apart than on the conv-only nets; the engine must not be further from the
        # CPU golden than the reference's own GPU arithmetic is
        assert de.max() <= dp.max() + 1e-3 and de.mean() <= 1.5 * dp.mean() + 1e-5


@pytest.mark.parametrize('key,scale,shape,crop', [('a4', 4, (3, 150, 300), 0), ('a2', 2, (2, 97, 127), 0), ('a2', 2, (1, 40, 126), 0), ('a2', 2, (3, 33, 253), 0),
                                                  ('dn_lite15', 1, (3, 260, 380), 0), ('a3', 3, (1, 19, 500), 0), ('a4', 4, (3, 300, 200), 96)])
def test_fused_residual_block_is_bit_identical_to_its_two_launch_form(engine, key, scale, shape, crop):
    """arsb_pair_kernel (conv_1 + PReLU + conv_2 + x scale + skip with the intermediate rows in shared memory, 126-px strips)
    against two launches of conv3x3_pair_trunk_kernel (128-px strips): same rounding points, same accumulation order ->
    the same bits, on widths around the strip boundaries (126, 127, 252, 253), several row segments, 1-3 planes, tiled plans"""
    from moephoto_b200 import runDN, imageProcess as IP
    from moephoto_b200.config import config
    try:
        if key.startswith('dn'):
            config.freeMemOverride = int(178 * 2 ** 30 * .9)
            opt = runDN.getOpt({'model': 'lite15'}, weights=H.load_weights(key))
        else:
            opt = _sr_opt(key, scale, crop)
        x = torch.rand(shape, generator=torch.Generator().manual_seed(sum(shape))).half().cuda()
        y = IP.doCrop(opt, x)                                # mid rows in tensor memory, conv_2 a .ts MMA (the default)
        assert torch.isfinite(y).all()
        for flags in (dict(no_arsb=True), dict(arsb_smem_mid=True), dict(arsb_solo=True)):
            engine.set_conv_path(**flags)
            try:
                y2 = IP.doCrop(opt, x)
            finally:
                engine.set_conv_path()
            assert torch.equal(y, y2), flags
    finally:
        config.freeMemOverride = None


@pytest.mark.parametrize('name', ['dn15_tiled', 'dn5_single_rgba', 'lite2_tiled', 'lite8_single'])
def test_skipping_the_zero_k_step_of_48_filter_models_changes_no_bit(engine, name):
    """NetDN and MoeNet_lite2 have 48 filters zero-padded to 64: channels 48..63 of every activation are exactly 0, so the kernels
    skip the fourth K step (16 channels) of every tap (ConvParams::ksteps).  `full_k` issues it anyway: same bits"""
    c = H.load_case(name)
    y = H.run_case_engine(c)
    engine.set_conv_path(full_k=True)
    try:
        y4 = H.run_case_engine(c)
    finally:
        engine.set_conv_path()
    assert np.array_equal(y, y4)


def test_two_streams_share_an_engine(engine):
    """every pair-kernel launch draws its items from its own counter block and the FRM scratch lives in the caller's
    workspace, so two plans may run concurrently on two streams of one engine (two workspaces): results equal the serial ones"""
    from moephoto_b200 import imageProcess as IP
    from moephoto_b200.config import config
    import ctypes
    oa, ol = _sr_opt('a4', 4, crop=96), _sr_opt('lite2', 2, crop=64)
    try:
        g = torch.Generator().manual_seed(31)
        xa, xl = torch.rand(3, 300, 520, generator=g).half().cuda(), torch.rand(3, 200, 300, generator=g).half().cuda()
        want_a, want_l = IP.doCrop(oa, xa), IP.doCrop(ol, xl)
        torch.cuda.synchronize()
        need = lambda o, x: engine.lib.moe_plan_workspace_bytes(o.modelCached.handle, 3, ctypes.byref(o.plan.c), 0, o.plan.out_h)
        wa = torch.empty(need(oa, xa), dtype=torch.uint8, device='cuda')
        wl = torch.empty(need(ol, xl), dtype=torch.uint8, device='cuda')
        s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
        for it in range(3):
            got_a, got_l = torch.full_like(want_a, float('nan')), torch.full_like(want_l, float('nan'))
            torch.cuda.synchronize()
            with torch.cuda.stream(s1):
                for _ in range(2):
                    IP.run_plan(oa.modelCached, xa, oa.plan, got_a, workspace=wa)
            with torch.cuda.stream(s2):
                for _ in range(4):
                    IP.run_plan(ol.modelCached, xl, ol.plan, got_l, workspace=wl)
            torch.cuda.synchronize()
            da, dl = (got_a.float() - want_a.float()).abs(), (got_l.float() - want_l.float()).abs()
            assert torch.equal(got_a, want_a), ('a4 on stream 1', float(da.max()), int((da > 0).sum()), [int(v) for v in (da > 0).nonzero()[0]])
            assert torch.equal(got_l, want_l), ('lite2 on stream 2', float(dl.max()), int((dl > 0).sum()), [int(v) for v in (dl > 0).nonzero()[0]])
    finally:
        config.freeMemOverride = None


@pytest.mark.parametrize('channels', [1, 4])
def test_host_buffer_entry_point_grey_and_rgba(engine, channels):
    """moe_enhance_host_c: a grey frame, and an RGBA frame whose alpha plane the SR model upscales like a colour plane
    (runSR.py:39-40) while the DN model passes it through (_RGBFilter, imageProcess.py:370-377) — equal to the API path"""
    from moephoto_b200 import _lib, runDN, imageProcess as IP
    from moephoto_b200.config import config
    opt = _sr_opt('a2', 2, crop=64, ram=int(4e9))
    try:
        img = np.random.default_rng(40 + channels).integers(0, 256, (70, 90, channels), dtype=np.uint8)
        want = IP.toOutput(8)(IP.doCrop(opt, IP.toTorch(8)(img)))
        out = np.empty((140, 180, channels), dtype=np.uint8)
        _lib.check(engine.lib.moe_enhance_host_c(opt.modelCached.handle, img.ctypes.data_as(ctypes.c_void_p), 8, channels, ctypes.byref(opt.plan.c),
                                                 out.ctypes.data_as(ctypes.c_void_p), 8, None))
        assert np.array_equal(out, want)
        if channels == 4:
            config.freeMemOverride, config.crop_dn = int(4e9), 48
            odn = runDN.getOpt({'model': 'lite15'}, weights=H.load_weights('dn_lite15'))
            want = IP.toOutput(8)(IP.RGBFilter(odn)(IP.toTorch(8)(img)))
            out = np.empty((70, 90, 4), dtype=np.uint8)
            _lib.check(engine.lib.moe_enhance_host_c(odn.modelCached.handle, img.ctypes.data_as(ctypes.c_void_p), 8, 4, ctypes.byref(odn.plan.c),
                                                     out.ctypes.data_as(ctypes.c_void_p), 8, None))
            assert np.array_equal(out, want)          # (alpha itself goes x/255 -> fp16 -> x256: 254 comes back as 255, in the reference too)
    finally:
        config.freeMemOverride, config.crop_dn = None, 'auto'


def test_video_pipe_loop_equals_the_reference_frame_loop(engine):
    """moephoto_b200.video.pipe_loop sits where the reference's loop sits (video.py:339-360): raw bgr48le bytes from a pipe
    read() -> batches of frames through DN -> SR on the engine -> bytes to a pipe write().  Against the per-frame chain of the
    reference's steps (toNumPy -> toTorch(16) -> RGBFilter -> sr -> toOutput(16) -> BGR -> toBuffer), byte for byte, with a
    frame count that is not a multiple of the batch and a short read ending the stream"""
    import io
    from moephoto_b200 import runSR, runDN, imageProcess as IP, video
    from moephoto_b200.config import config
    rng = np.random.default_rng(8)
    h, w, n = 40, 56, 7
    raw_frames = [rng.integers(0, 65536, (h, w, 3), dtype=np.uint16).tobytes() for _ in range(n)]
    config.freeMemOverride, config.crop_dn, config.crop_sr = int(4e9), 32, 32
    try:
        odn = runDN.getOpt({'model': 'lite15'}, weights=H.load_weights('dn_lite15'))
        osr = runSR.getOpt({'model': 'a', 'scale': 2}, weights=H.load_weights('a2'))
        want = b''
        for raw in raw_frames:                                         # the reference's loop: one frame per iteration
            im = IP.toNumPy(16)((raw, h, w))
            x = IP.toTorch(16, swapRB=True)(im)
            want += IP.toBuffer(16)(IP.toOutput(16, swapRB=True)(runSR.sr(osr)(IP.RGBFilter(odn)(x))))
        pipe_in = io.BytesIO(b''.join(raw_frames) + b'\x00' * 100)      # a truncated last read ends the stream
        pipe_out = io.BytesIO()
        done = video.pipe_loop(pipe_in.read, pipe_out.write, h, w, [odn, osr], bit_depth=16, swap_rb=True, batch=3)
        assert done == n and pipe_out.getvalue() == want
        assert video.pipe_loop(io.BytesIO(b'').read, pipe_out.write, h, w, [odn, osr]) == 0
    finally:
        config.freeMemOverride, config.crop_dn, config.crop_sr = None, 'auto', 'auto'


def test_engine_check_reports_no_timeout_in_normal_operation(engine):
    """the kernels' mbarrier waits give up after a time-out by raising a device flag (no __trap: DESIGN.md §5, include/moephoto_b200.h);
    moe_engine_check is how the host sees it.  In normal operation, even with the time-out shrunk to 50 ms, it never fires."""
    from moephoto_b200 import _lib, imageProcess as IP
    from moephoto_b200.config import config
    opt = _sr_opt('a4', 4, crop=96)
    try:
        x = torch.rand(3, 200, 300, generator=torch.Generator().manual_seed(1)).half().cuda()
        want = IP.doCrop(opt, x).clone()
        _lib.check(engine.lib.moe_engine_debug_timeout(engine.handle, 50000000))
        got = IP.doCrop(opt, x)
        assert engine.lib.moe_engine_check(engine.handle, None) == _lib.MOE_OK
        assert torch.equal(got, want)
    finally:
        engine.lib.moe_engine_debug_timeout(engine.handle, 0)
        config.freeMemOverride = None
