"""The oracle against the committed golden vectors (outputs of the unmodified reference).
  `.ref`   : the reference's CPU fp32 output.  The fixtures hold the checkpoints rounded to fp16 — the weights the
             reference's GPU path computes with — so the residual of the fp32 oracle is the weight rounding alone: PSNR > 70 dB.
  `.ref16` : the reference in its GPU fp16 configuration (model.half(), half tensors; executed on CPU).  The oracle's
             mode='ref16cpu' has the same rounding points (mode='ref16', the engine's contract, differs from it in ONE place:
             on the GPU a convolution's bias add is a separate aten op and rounds separately — oracle/net.py::_biased,
             profiles/r02_cudnn_rounding_probe.log); what remains is the fp32 summation order inside conv2d, which
             flips an fp16 rounding here and there and propagates.  THE BAR (also the engine's, tests/test_gpu_engine.py):
               smooth images       max-abs <= 1e-3  (one fp16 ulp at 1.0 = 9.77e-4; measured exactly that), PSNR >= 75 dB
               uniform white noise max-abs <= 2e-3  (two ulps), at most 0.2 % of the pixels beyond 1e-3, PSNR >= 68 dB
             — measured for oracle vs reference: 78-84 dB / 70-72 dB (profiles/r02_parity_oracle_vs_reference.txt)."""
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
  if H.is_white_noise(c):     # fp16 weights on white noise: the reference's own fp16 configuration is 60-69 dB from its fp32 output here
    assert H.psnr(y, c['ref']) > 60.0 and d.max() < 3e-2
  else:
    assert H.psnr(y, c['ref']) > 70.0
    assert d.max() < 1.5e-2 and np.quantile(d, .999) < 2e-3
  if c['alpha'] is not None:
    assert np.array_equal(y[3], c['alpha'])          # alpha bypasses the denoiser


@pytest.mark.parametrize('name', H.case_names())
def test_oracle_ref16_reproduces_the_reference_fp16_golden(name):
  """every case, ensemble and p* weights included, against the reference's own fp16-configuration output"""
  c = H.load_case(name)
  y = H.run_case_oracle(c, mode='ref16cpu')
  H.assert_ref16_bar(y, c)
  if c['alpha'] is not None:
    assert np.array_equal(y[3], c['alpha'].astype(np.float16).astype(np.float32))


@pytest.mark.parametrize('name', ['a2_tiled', 'a4_single', 'dn15_tiled', 'lite4_single', 'a4_rand'])
def test_gpu_and_cpu_bias_semantics_differ_only_where_a_convolution_has_a_bias(name):
  """mode 'ref16' (GPU: the bias add is its own op) vs 'ref16cpu': identical for NetDN (no biased convolution), one ulp on
  a fraction of the pixels otherwise — and still inside the cross-platform bar against the CPU-executed golden"""
  c = H.load_case(name)
  g, k = H.run_case_oracle(c, mode='ref16'), H.run_case_oracle(c, mode='ref16cpu')
  if c['kind'] == 'dn':
    assert np.array_equal(g, k)
  else:
    assert not np.array_equal(g, k)
  H.assert_cross_platform_bar(g, c)


@pytest.mark.parametrize('name', ['a2_tiled', 'a4_single', 'dn15_tiled'])
def test_round1_f16io_mode_is_further_from_the_reference_than_ref16(name):
  """one rounding per stored tensor (round 1) vs the reference's per-op roundings: both within the fp16 gap of the fp32
  output, ref16 several dB closer to what the reference's half model produces"""
  c = H.load_case(name)
  y1, y2 = H.run_case_oracle(c, mode='f16io'), H.run_case_oracle(c, mode='ref16cpu')
  assert H.psnr(y1, c['ref']) > 66.0                  # reference fp16-vs-fp32 is 71-73 dB on 128x128
  assert H.psnr(y2, c['ref16']) > H.psnr(y1, c['ref16']) + 3.0


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
  sd = H.load_weights('lite4')
  assert np.abs(N.forward_torch(sd, x) - N.forward(sd, x)).max() < 2e-5


def test_forward_torch_in_half_on_cpu_is_the_ref16cpu_mode():
  """the PyTorch port run in half (every op rounds, as in the reference's half model) against mode='ref16'"""
  from oracle import net as N
  x = np.random.default_rng(4).random((2, 1, 20, 28)).astype(np.float16).astype(np.float32)
  for key in ('a2', 'dn_lite15', 'lite2'):
    sd = H.load_weights(key)
    d = np.abs(N.forward_torch(sd, x, dtype='float16') - N.forward(sd, x, mode='ref16cpu'))     # on the CPU the bias is inside the conv
    assert d.max() <= 2e-3 and (d > 1e-3).mean() <= 2e-3, key
