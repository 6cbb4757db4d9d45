"""Generates the committed golden fixtures from the UNMODIFIED reference, in the build container only
(needs /root/reference).  Run:  python tests/golden/make_golden.py

Writes next to this file:
  weights_<arch>.npz   fp16 copies of the reference checkpoints the BASELINE configs name
                       (exactly `model.half()` — what the reference's GPU fp16 path computes with;
                       imageProcess.py:311-318 castModel).  Keys = checkpoint keys.
  cases.npz            small cases.  For each: uint8 HWC input image, `<case>.ref` = the reference's fp32 CPU output
                       (runSR.sr / imageProcess.RGBFilter), `<case>.ref16` = the SAME call with the reference switched
                       to its GPU fp16 configuration (model.half(), half tensors, half blend ramp — every aten op rounds
                       to fp16 as on the GPU; executed on CPU, oracle/refharness.py::gpu_fp16_config), its tile list
                       (opt.iterClip()) and the free-memory figure the plan was made with.
  bands.npz            large cases (a 512x1024 a4 frame in 18 tiles; a 96x3840 a4 frame in the 4 column strips of the
                       reference's 4K plan): the input image and, of the fp16-configuration output, every seam band
                       (full length, 40 px across the seam) plus interior windows — the full outputs are 50 / 35 MB.
The reference ships no golden vectors; these are outputs of the reference itself.
"""
import os
import sys
import json
import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, '..', '..')))
from oracle import refharness as R   # noqa: E402

CKPT = {'a2': 'a2/model_new.pth', 'a3': 'a3/model_new.pth', 'a4': 'a4/model_new.pth', 'p4': 'p4/model_new.pth',
        'dn_lite15': 'dn_lite15/model_new.pth', 'dn_lite5': 'dn_lite5/model_new.pth',
        'lite2': 'lite/model.pth', 'lite4': 'lite/model_4.pth', 'lite8': 'lite/model_8.pth'}
RAM = 64295018496          # the free-memory figure of the round-1 goldens (a 62 GB host), pinned so re-runs reproduce them
RAM_4STRIP = int(6.54e9)   # with this budget prepare() cuts a 96x3840 a4 frame into the column strips of the 4K plan


def smooth_noise_u8(h, w, seed):
  """SURVEY.md §8d recipe: bicubic-upsampled low-res noise + 0.03 randn, quantised to 8 bit (HWC)."""
  g = torch.Generator().manual_seed(seed)
  lo = torch.rand(1, 3, max(2, (h + 7) // 8), max(2, (w + 7) // 8), generator=g)
  x = torch.nn.functional.interpolate(lo, size=(h, w), mode='bicubic', align_corners=False).clamp(0, 1)[0]
  x = (x + 0.03 * torch.randn(3, h, w, generator=g)).clamp(0, 1)
  return (x * 255).round().to(torch.uint8).permute(1, 2, 0).contiguous().numpy()


def uniform_u8(h, w, seed):
  """uniform white noise, the worst case for seams and for fp16 rounding flips"""
  return np.random.default_rng(seed).integers(0, 256, (h, w, 3), dtype=np.uint8)


CASES = [
  # name, kind, arg, (H, W), crop, seed, options
  ('a2_single', 'sr', 2, (64, 96), 0, 1, {}),
  ('a2_unaligned', 'sr', 2, (45, 70), 0, 2, {}),          # reflect-pad-to-8 on both axes
  ('a2_tiled', 'sr', 2, (72, 100), 48, 3, {}),
  ('a3_tiled', 'sr', 3, (64, 88), 48, 4, {}),              # pad 9
  ('a4_tiled', 'sr', 4, (64, 96), 48, 5, {}),
  ('a4_single', 'sr', 4, (40, 56), 0, 6, {}),
  ('dn15_tiled', 'dn', 'lite15', (72, 96), 48, 7, {}),     # pad 7
  ('dn5_single_rgba', 'dn', 'lite5', (48, 64), 0, 8, {}),  # alpha bypass (4th plane appended below)
  # MoeNet_lite2 (runSR.py:21-23): FRM's global mean makes the result depend on the tile, so tiled cases matter
  ('lite2_tiled', 'sr:lite', 2, (72, 100), 48, 9, {}),
  ('lite4_single', 'sr:lite', 4, (40, 56), 0, 10, {}),
  ('lite8_single', 'sr:lite', 8, (32, 40), 0, 11, {}),
  # round 2: uniform white noise through every net (2x2 tiles), the dihedral ensemble, the p* weights
  ('a2_rand', 'sr', 2, (56, 72), 48, 12, dict(uniform=True)),
  ('a3_rand', 'sr', 3, (56, 72), 48, 13, dict(uniform=True)),
  ('a4_rand', 'sr', 4, (56, 72), 48, 14, dict(uniform=True)),
  ('dn15_rand', 'dn', 'lite15', (56, 72), 48, 15, dict(uniform=True)),
  ('lite4_rand', 'sr:lite', 4, (56, 72), 48, 16, dict(uniform=True)),
  ('a2_ens3', 'sr', 2, (48, 72), 48, 17, dict(ensemble=3)),          # runSR.py:26, imageProcess.py:564-572
  ('a2_ens7', 'sr', 2, (48, 72), 48, 18, dict(ensemble=7)),
  ('p4_single', 'sr:p', 4, (40, 56), 0, 19, {}),                      # runSR.py:14-16: same nets, other checkpoints
]

BIG = [
  # name, scale, (H, W), crop, ram, seed
  ('a4_big', 4, (512, 1024), 192, RAM, 31),                # 3 x 6 tiles
  ('a4_4strip', 4, (96, 3840), 0, RAM_4STRIP, 32),         # columns 0-968, 963-1931, 1921-2889, 2880-3840 as for 3840x2160
]


def seam_windows(plan, scale, pad, out_h, out_w):
  """(y0,y1,x0,x1) windows of the canvas: a 40-px band over every seam (8 px before its first blended row/column,
  the pad*scale blended ones, the rest after) and four interior squares"""
  wins, psc = [], pad * scale
  for top in sorted({t[0] for t in plan if t[0] > 0}):
    y = top * scale
    wins.append((max(0, y - 8), min(out_h, y + 32), 0, out_w))
  for left in sorted({t[2] for t in plan if t[2] > 0}):
    x = left * scale
    wins.append((0, out_h, max(0, x - 8), min(out_w, x + 32)))
  s = min(128, out_h // 2)
  for fy, fx in ((.2, .2), (.2, .7), (.6, .4), (.7, .8)):
    y, x = int(fy * (out_h - s)), int(fx * (out_w - s))
    wins.append((y, y + s, x, x + s))
  return wins


def pin_free_memory(ref, ram):
  ref['config'].calcFreeMem = (lambda r: (lambda *a, **k: r))(int(ram))


def main():
  ref = R.load()
  for name, rel in CKPT.items():
    sd = R.state_dict(rel)
    np.savez(os.path.join(HERE, 'weights_%s.npz' % name),
             **{k: v.numpy().astype(np.float16) for k, v in sd.items()})
  out, meta = {}, {}
  to_tensor = lambda im: torch.from_numpy(im).permute(2, 0, 1).float() / 255   # == torchvision to_tensor
  for name, kind, arg, (h, w), crop, seed, o in CASES:
    img = uniform_u8(h, w, seed) if o.get('uniform') else smooth_noise_u8(h, w, seed)
    x = to_tensor(img)
    pin_free_memory(ref, RAM)
    model = 'a'
    if kind.startswith('sr:'):
      kind, model = kind.split(':')
    ys = []
    for half in (False, True):
      if kind == 'sr':
        y, plan, opt = R.run_sr(x, arg, crop=crop, model=model, ensemble=o.get('ensemble', 0), half=half)
      else:
        if name.endswith('rgba') and x.shape[0] == 3:
          g = torch.Generator().manual_seed(seed + 100)
          x = torch.cat([x, torch.rand(1, h, w, generator=g)], 0)
        y, plan, opt = R.run_dn(x, arg, crop=crop, half=half)
      ys.append((y, [[int(v) for v in t] for t in plan], float(opt.ramCoef)))
    assert ys[0][1] == ys[1][1], 'the fp32 and fp16 configurations must cut the same tiles'
    assert ys[1][0].dtype == torch.half
    out[name + '.img'] = img
    if x.shape[0] == 4:
      out[name + '.alpha'] = x[3].numpy()
    out[name + '.ref'] = ys[0][0].numpy().astype(np.float32)
    out[name + '.ref16'] = ys[1][0].numpy()
    meta[name] = dict(kind=kind, arg=arg, model=model, crop=crop, ram=RAM, pad=int(opt.padding), scale=int(opt.scale),
                      ram_coef=ys[0][2], ram_coef_gpu=ys[1][2], ensemble=int(o.get('ensemble', 0)), tiles=ys[0][1])
    d = (ys[1][0].float() - ys[0][0]).abs()
    print(name, tuple(y.shape), len(plan), 'tiles; reference fp16 vs fp32 max %.2e' % d.max().item())
  out['meta'] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
  np.savez_compressed(os.path.join(HERE, 'cases.npz'), **out)

  out, meta = {}, {}
  for name, scale, (h, w), crop, ram, seed in BIG:
    img = smooth_noise_u8(h, w, seed)
    pin_free_memory(ref, ram)
    y, plan, opt = R.run_sr(to_tensor(img), scale, crop=crop, half=True)
    y = y.numpy()
    wins = seam_windows(plan, scale, int(opt.padding), y.shape[-2], y.shape[-1])
    out[name + '.img'] = img
    for i, (y0, y1, x0, x1) in enumerate(wins):
      out['%s.win%d' % (name, i)] = np.ascontiguousarray(y[:, y0:y1, x0:x1])
    meta[name] = dict(kind='sr', arg=scale, model='a', crop=crop, ram=int(ram), pad=int(opt.padding), scale=scale,
                      ram_coef_gpu=float(opt.ramCoef), tiles=[[int(v) for v in t] for t in plan], windows=[list(map(int, v)) for v in wins])
    print(name, tuple(y.shape), len(plan), 'tiles', len(wins), 'windows')
  out['meta'] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
  np.savez_compressed(os.path.join(HERE, 'bands.npz'), **out)


if __name__ == '__main__':
  main()
