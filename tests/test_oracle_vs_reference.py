"""Pins the oracle (oracle/tiling.py, oracle/net.py) and the product's host planner against the
UNMODIFIED reference imported on CPU (oracle/refharness.py).  Only runs where /root/reference exists
(the build container); the GPU box relies on the committed goldens generated from the same code."""
import numpy as np
import pytest
import torch

from oracle import refharness as R, net as N, tiling as T

pytestmark = pytest.mark.skipif(not R.available(), reason='reference tree not present')


@pytest.fixture(scope='module')
def ref():
  return R.load()


def _ref_plan(ref, shape, ram, ram_coef, pad, sc, align, crop):
  ip = ref['imageProcess']
  opt = ip.Option()
  opt.fixChannel, opt.ramCoef = 0, ram_coef
  it, *_ = ip.prepare(shape, ram, opt, pad, sc, align, crop)
  return [tuple(int(v) for v in t) for t in it()]


SWEEP = [(h, w, crop) for h in (40, 45, 64, 97, 130, 256, 301) for w in (40, 70, 96, 128, 233, 512) for crop in (0, 48, 64, 120)]


@pytest.mark.parametrize('pad,sc,coef', [(5, 2, .9 / 2473.), (9, 3, .9 / 6120.), (5, 4, .9 / 7029.7), (7, 1, .95 / 1253.4)])
def test_tile_plans_match_reference(ref, pad, sc, coef):
  from moephoto_b200 import imageProcess as IP
  opt = IP.Option()
  opt.fixChannel, opt.ramCoef = 0, coef
  n = 0
  for ram in (3e8, 4e9, 1.6e11):
    for h, w, crop in SWEEP:
      shape = (3, h, w)
      try:
        want = _ref_plan(ref, shape, ram, coef, pad, sc, 8, crop)
      except MemoryError:
        with pytest.raises(MemoryError):
          T.make_plan(shape, ram, coef, pad, sc, 8, crop)
        with pytest.raises(MemoryError):
          IP.makePlan(shape, ram, opt, pad, sc, 8, crop)
        continue
      assert T.make_plan(shape, ram, coef, pad, sc, 8, crop).tiles == want, (shape, ram, crop)
      assert IP.makePlan(shape, ram, opt, pad, sc, 8, crop).tiles == want, (shape, ram, crop)
      n += 1
  assert n > 100


def test_big_frames_plan(ref):
  """the BASELINE shapes with the memory of a B200: 1 tile for 1080p a2, 4 strips for 4K a4"""
  from moephoto_b200 import imageProcess as IP
  ram = int(178 * 2 ** 30 * .9)
  for shape, pad, sc, coef, crop in (((3, 1080, 1920), 5, 2, .9 / 2473., 0), ((3, 2160, 3840), 5, 4, .9 / 7029.7, 0),
                                     ((3, 2160, 3840), 5, 4, .9 / 7029.7, 512), ((3, 2160, 3840), 9, 3, .9 / 6120., 0),
                                     ((48, 1080, 1920), 7, 1, .95 / 1253.4, 0)):
    want = _ref_plan(ref, shape, ram, coef, pad, sc, 8, crop)
    opt = IP.Option()
    opt.fixChannel, opt.ramCoef = 0, coef
    assert T.make_plan(shape, ram, coef, pad, sc, 8, crop).tiles == want
    assert IP.makePlan(shape, ram, opt, pad, sc, 8, crop).tiles == want


def _fake_net_np(sc):
  def f(a):
    n, _, h, w = a.shape
    up = a.repeat(sc, axis=2).repeat(sc, axis=3)
    ry = (np.arange(h * sc, dtype=np.float32) / np.float32(h * sc))[:, None] * np.float32(.01)
    rx = (np.arange(w * sc, dtype=np.float32) / np.float32(w * sc))[None, :] * np.float32(.02)
    return (up + ry + rx).astype(np.float32)
  return f


def _fake_net_torch(sc):
  def f(a):
    n, _, h, w = a.shape
    up = a.repeat_interleave(sc, 2).repeat_interleave(sc, 3)
    ry = (torch.arange(h * sc, dtype=torch.float32) / (h * sc))[:, None] * .01
    rx = (torch.arange(w * sc, dtype=torch.float32) / (w * sc))[None, :] * .02
    return up + ry + rx
  return f


@pytest.mark.parametrize('h,w,crop,pad,sc', [(64, 96, 48, 5, 2), (45, 70, 0, 5, 2), (100, 130, 48, 9, 3), (72, 100, 40, 7, 1),
                                             (90, 61, 48, 5, 4), (33, 250, 64, 5, 2), (7, 300, 64, 5, 2)])
def test_stitching_matches_reference_bit_for_bit(ref, h, w, crop, pad, sc):
  """doCrop with a tile-position-dependent fake network: padImage, unpad, both blends and the
  bottom-right aligned store must reproduce the reference's canvas exactly (fp32)."""
  ip = ref['imageProcess']
  g = torch.Generator().manual_seed(h * 1000 + w)
  x = torch.rand(3, h, w, generator=g)
  ram, coef = 4e9, .9 / 2473.
  opt = ip.Option()
  opt.fixChannel, opt.ramCoef, opt.padding, opt.scale, opt.cropsize = 0, coef, pad, sc, crop
  opt.squeeze, opt.unsqueeze = (lambda t: t.squeeze(1)), (lambda t: t.unsqueeze(1))
  opt.modelCached = _fake_net_torch(sc)
  ref['config'].calcFreeMem = lambda *a, **k: int(ram)
  want = ip.doCrop(opt, x).numpy()
  plan = T.make_plan((3, h, w), int(ram), coef, pad, sc, 8, crop)
  # the sigmoid ramp comes from torch's vectorised libm: equal to ours within one fp32 ulp; the stitching
  # itself (given the ramp) must be bit-exact
  assert np.abs(T.blend_ramp(plan.pad_sc) - opt.blend.numpy().reshape(-1)).max() <= 1.2e-7
  got = T.do_crop(_fake_net_np(sc), x.numpy(), plan, ramp=opt.blend.numpy().reshape(-1))
  assert got.shape == want.shape
  assert np.array_equal(got, want)
  assert np.abs(T.do_crop(_fake_net_np(sc), x.numpy(), plan) - want).max() <= 2.4e-7


@pytest.mark.parametrize('key,ckpt', [('net2x', 'a2'), ('net3x', 'a3'), ('net4x', 'a4'), ('netdn', 'dn_lite15'), ('net2x', 'p2')])
def test_network_forward_matches_reference(ref, key, ckpt):
  sd_t = R.state_dict(ckpt + '/model_new.pth')
  ctor = {'net2x': 'Net2x', 'net3x': 'Net3x', 'net4x': 'Net4x', 'netdn': 'NetDN'}[key]
  model = getattr(ref['models'], ctor)()
  model.load_state_dict(sd_t)
  model.eval()
  g = torch.Generator().manual_seed(7)
  x = torch.rand(2, 1, 24, 40, generator=g)
  with torch.no_grad():
    want = model(x)[-1].numpy()
  sd = N.to_numpy_state(sd_t)
  assert N.arch_of_state_dict(sd) == key
  for backend in ('c', 'torch'):
    got = N.forward(sd, x.numpy(), backend=backend)
    assert got.shape == want.shape
    assert np.abs(got - want).max() < 2e-5, backend


def test_frame_conversions_match_reference(ref):
  ip = ref['imageProcess']
  rng = np.random.default_rng(0)
  img8 = rng.integers(0, 256, (9, 13, 3), dtype=np.uint8)
  a = ip.toTorch(8, torch.float, 'cpu')(img8).numpy()
  assert np.array_equal(a, T.to_planar(img8, 8))
  img16 = rng.integers(0, 65536, (9, 13, 3), dtype=np.uint16)
  b = ip.toTorch(16, torch.float, 'cpu')(img16.astype(np.float32)).numpy()
  assert np.array_equal(b, T.to_planar(img16, 16))
  y = torch.from_numpy(rng.uniform(-.1, 1.1, (3, 9, 13)).astype(np.float32))
  for bits in (8, 16):
    want = ip.toOutput(bits)(ip.toFloat(y))
    got = T.to_output(y.numpy(), bits)
    assert np.array_equal(want.astype(np.int64), got.astype(np.int64))


@pytest.mark.parametrize('upscale,ckpt', [(2, 'lite/model.pth'), (4, 'lite/model_4.pth'), (8, 'lite/model_8.pth')])
def test_lite_network_forward_matches_reference(ref, upscale, ckpt):
  """MoeNet_lite2.Net (runSR.py:21-23): 1x1 convs, LB blocks with the FRM gate, PixelShuffle(2) stages"""
  import importlib
  lite = importlib.import_module('MoeNet_lite2')
  sd_t = R.state_dict(ckpt)
  model = lite.Net(upscale=upscale)
  model.load_state_dict(sd_t)
  model.eval()
  x = torch.rand(2, 1, 24, 40, generator=torch.Generator().manual_seed(3))
  with torch.no_grad():
    want = model(x)[-1].numpy()
  sd = N.to_numpy_state(sd_t)
  assert N.arch_of_state_dict(sd) == 'lite'
  got = N.forward(sd, x.numpy())
  assert got.shape == want.shape and np.abs(got - want).max() < 2e-5


@pytest.mark.parametrize('ckpt,getopt,key', [('a4/model_new.pth', ('runSR', {'model': 'a', 'scale': 4}), 'SRa4'),
                                             ('a3/model_new.pth', ('runSR', {'model': 'a', 'scale': 3}), 'SRa3'),
                                             ('dn_lite15/model_new.pth', ('runDN', {'model': 'lite15'}), 'DNlite15'),
                                             ('lite/model.pth', ('runSR', {'model': 'lite', 'scale': 2}), 'SRlite2')])
def test_ref16_mode_matches_the_reference_run_in_half(ref, ckpt, getopt, key):
  """mode='ref16cpu' restates the rounding points of `model.half()` on half tensors (imageProcess.py:309-318 castModel):
  the UNMODIFIED reference network, cast to half and run live on CPU, against the numpy/C restatement on white noise —
  the worst case.  Identical rounding points, different fp32 summation order inside conv2d: <= 2 fp16 ulps at 1.0,
  (almost) no pixel beyond one ulp, and far closer than the round-1 contract (one rounding per stored tensor)."""
  x = torch.rand(2, 40, 56, generator=torch.Generator().manual_seed(11)).half()
  with R.gpu_fp16_config():
    getattr(ref[getopt[0]], 'getOpt')(dict(getopt[1]))
    net = R.bare_net(key)
    with torch.no_grad():
      want = net(x.unsqueeze(1))[-1]
    assert want.dtype == torch.half
    want = want.float().numpy()
  sd = N.to_numpy_state(R.state_dict(ckpt))
  xn = x.float().numpy()[:, None]
  got, old = N.forward(sd, xn, mode='ref16cpu'), N.forward(sd, xn, mode='f16io')
  d = np.abs(got - want)
  assert d.max() <= 2e-3 and (d > 1e-3).mean() <= 2e-3
  assert d.mean() < 0.8 * np.abs(old - want).mean()


def test_ensemble_restatement_matches_reference(ref):
  """T.ensemble (imageProcess.py:558-572 + runSR.py:26) against the reference's own sr() with ensemble = 1..7, fp32, on a
  non-square tiled image (the transposed passes run on the plan of the transposed shape)"""
  x = torch.rand(3, 44, 70, generator=torch.Generator().manual_seed(21))
  sd = N.to_numpy_state(R.state_dict('a2/model_new.pth'))
  ram = 4e9
  ref['config'].calcFreeMem = lambda *a, **k: int(ram)
  for k in (1, 4, 7):
    want, tiles, opt = R.run_sr(x, 2, crop=48, ensemble=k)
    coef = float(opt.ramCoef)
    plan = T.make_plan((3, 44, 70), int(ram), coef, 5, 2, 8, 48)
    plan_t = T.make_plan((3, 70, 44), int(ram), coef, 5, 2, 8, 48)
    assert plan.tiles == [tuple(int(v) for v in t) for t in tiles]
    got = T.ensemble(lambda a: N.forward(sd, a), x.numpy(), plan, plan_t, k)
    assert got.shape == tuple(want.shape) and np.abs(got - want.numpy()).max() < 2e-5, k
