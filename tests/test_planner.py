"""Host logic of the product (tile planner, ramp, Option plumbing) against the oracle and the goldens.
Runs anywhere (no reference tree, no GPU)."""
import numpy as np
import pytest

import helpers as H
from oracle import tiling as T
from moephoto_b200 import imageProcess as IP


def _opt(coef):
  o = IP.Option()
  o.fixChannel, o.ramCoef = 0, coef
  return o


@pytest.mark.parametrize('pad,sc,coef', [(5, 2, .9 / 2473.), (9, 3, .9 / 6120.), (5, 4, .9 / 7029.7), (7, 1, .95 / 1253.4)])
def test_planner_equals_oracle_over_a_sweep(pad, sc, coef):
  rng = np.random.default_rng(pad * 10 + sc)
  for _ in range(300):
    h, w = int(rng.integers(1, 700)), int(rng.integers(1, 900))
    crop = int(rng.choice([0, 0, 40, 48, 64, 100, 128, 256, 512]))
    ram = float(rng.choice([2e8, 1e9, 8e9, 1.6e11]))
    planes = int(rng.choice([1, 3, 4, 48]))
    shape = (planes, h, w)
    try:
      want = T.make_plan(shape, ram, coef, pad, sc, 8, crop)
    except MemoryError:
      with pytest.raises(MemoryError):
        IP.makePlan(shape, ram, _opt(coef), pad, sc, 8, crop)
      continue
    got = IP.makePlan(shape, ram, _opt(coef), pad, sc, 8, crop)
    assert got.tiles == want.tiles, (shape, ram, crop)
    assert (got.pad_h, got.pad_w, got.out_h, got.out_w, got.pad_sc) == (want.pad_h, want.pad_w, want.out_h, want.out_w, want.pad_sc)


def test_planner_reproduces_the_golden_tile_lists():
  for name in H.case_names():
    c = H.load_case(name)
    h, w = c['img'].shape[:2]
    planes = 3
    got = IP.makePlan((planes, h, w), c['ram'], _opt(c['ram_coef']), c['pad'], c['scale'], 8, c['crop'])
    assert [list(t) for t in got.tiles] == c['tiles'], name


def test_ramp_is_the_fp16_sigmoid():
  for psc in (7, 10, 20, 27):
    a = IP.blendRamp(psc)
    b = T.blend_ramp(psc, np.float16).astype(np.float32)
    assert a.shape == (psc,)
    assert np.abs(a - b).max() <= 2 ** -11      # one fp16 ulp below 1
    assert np.all(np.diff(a) > 0) and abs(a[psc // 2] - (0.5 if psc % 2 == 0 else a[psc // 2])) < 1e-6


def test_plan_struct_roundtrip():
  p = IP.makePlan((3, 72, 100), 4e9, _opt(.9 / 2473.), 5, 2, 8, 48)
  assert p.c.n_tiles == len(p.tiles) == 6
  t = p.c.tiles[1]
  assert (t.top, t.bottom, t.left, t.right, t.top_t, t.left_t, t.bsc, t.rsc) == p.tiles[1]
  assert p.c.out_h == 144 and p.c.out_w == 200 and p.c.pad_sc == 10
  assert abs(p.c.ramp[5] - 0.5) < 1e-6


def test_prepareOpt_caches_like_the_reference(monkeypatch):
  """imageProcess.py:133-155: re-plan on first use, when the plane count changes, and after 29 cached calls"""
  calls = []
  monkeypatch.setattr(IP.config, 'calcFreeMem', lambda *a, **k: calls.append(1) or int(4e9))
  o = _opt(.9 / 2473.)
  o.padding, o.scale = 5, 2
  IP.prepareOpt(o, (3, 64, 96))
  assert len(calls) == 1 and o.outShape == [3, 128, 192] and len(list(o.iterClip())) == 1
  for _ in range(29):
    IP.prepareOpt(o, (3, 64, 96))
  assert len(calls) == 1
  IP.prepareOpt(o, (3, 64, 96))
  assert len(calls) == 2
  IP.prepareOpt(o, (4, 64, 96))
  assert len(calls) == 3


def test_getOpt_unknown_models():
  from moephoto_b200 import runSR, runDN
  assert runSR.getOpt({'model': 'nope', 'scale': 2}) is None
  assert runSR.getOpt({'model': 'a', 'scale': 8}) is None
  with pytest.raises(KeyError):
    runDN.getOpt({'model': 'nope'})
  assert set(runSR.mode_switch) == {'a2', 'a3', 'a4', 'p2', 'p3', 'p4', 'lite2', 'lite4', 'lite8'}
  assert runSR.mode_switch['a4'][0] == './model/a4/model_new.pth' and abs(runSR.mode_switch['a4'][2][2] - .9 / 7029.7) < 1e-12
  assert runDN.mode_switch['lite15'][3:] == (1, 7, 8)


def test_video_route_byte_conversions():
  """toNumPy / toBuffer (imageProcess.py:216-236): raw bgr48le bytes <-> HWC uint16 frames, bit-exact round trip"""
  rng = np.random.default_rng(1)
  frame = rng.integers(0, 65536, (6, 9, 3), dtype=np.uint16)
  raw = frame.tobytes()
  back = IP.toNumPy(16)((raw, 6, 9))
  assert back.dtype == np.uint16 and np.array_equal(back, frame)
  assert IP.toBuffer(16)(back) == raw and IP.toBuffer(16)(None) is None and IP.toNumPy(16)((b'', 6, 9)) is None
  f8 = rng.integers(0, 256, (4, 5, 3), dtype=np.uint8)
  assert IP.toBuffer(8)(IP.toNumPy(8)((f8.tobytes(), 4, 5))) == f8.tobytes()
