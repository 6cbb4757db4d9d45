"""The oracle against the committed golden vectors (outputs of the unmodified reference, CPU fp32).
The fixtures hold the checkpoints rounded to fp16 — the weights the reference's GPU path computes with —
so the residual here is the weight rounding alone (documented in DESIGN.md §numerics): PSNR > 70 dB."""
import numpy as np
import pytest

import helpers as H


@pytest.mark.parametrize('name', H.case_names())
def test_oracle_fp32_reproduces_golden(name):
  c = H.load_case(name)
  assert [tuple(t) for t in c['tiles']] == H.oracle_plan(c).tiles
  y = H.run_case_oracle(c, mode='fp32')
  assert y.shape == c['ref'].shape
  d = np.abs(y - c['ref'])
  assert H.psnr(y, c['ref']) > 70.0
  assert d.max() < 1.5e-2 and np.quantile(d, .999) < 2e-3
  if c['alpha'] is not None:
    assert np.array_equal(y[3], c['alpha'])          # alpha bypasses the denoiser


@pytest.mark.parametrize('name', ['a2_tiled', 'a4_single', 'dn15_tiled'])
def test_oracle_f16io_mode_stays_within_the_reference_fp16_gap(name):
  c = H.load_case(name)
  y = H.run_case_oracle(c, mode='f16io')
  assert H.psnr(y, c['ref']) > 66.0                   # reference fp16-vs-fp32 is 71-73 dB on 128x128


def test_conv_backends_agree():
  from oracle import net as N
  rng = np.random.default_rng(1)
  x = rng.standard_normal((2, 5, 9, 11)).astype(np.float32)
  w = rng.standard_normal((7, 5, 3, 3)).astype(np.float32)
  b = rng.standard_normal(7).astype(np.float32)
  a, t = N.conv3x3(x, w, b, 'c'), N.conv3x3(x, w, b, 'torch')
  assert np.abs(a - t).max() < 1e-4
  ps = N.pixel_shuffle(np.arange(2 * 8 * 3 * 4, dtype=np.float32).reshape(2, 8, 3, 4), 2)
  import torch
  assert np.array_equal(ps, torch.nn.functional.pixel_shuffle(torch.arange(2 * 8 * 3 * 4, dtype=torch.float32).reshape(2, 8, 3, 4), 2).numpy())


def test_forward_torch_equals_forward():
  from oracle import net as N
  sd = H.load_weights('a4')
  x = np.random.default_rng(3).random((2, 1, 20, 28)).astype(np.float32)
  assert np.abs(N.forward_torch(sd, x) - N.forward(sd, x)).max() < 2e-5
  sd = H.load_weights('dn_lite15')
  assert np.abs(N.forward_torch(sd, x) - N.forward(sd, x)).max() < 2e-5
