"""Generates the committed golden fixtures from the UNMODIFIED reference (CPU fp32), in the build
container only (needs /root/reference).  Run:  python tests/golden/make_golden.py

Writes next to this file:
  weights_<arch>.npz   fp16 copies of the reference checkpoints the BASELINE configs name
                       (exactly `model.half()` — what the reference's GPU fp16 path computes with;
                       imageProcess.py:311-318 castModel).  Keys = checkpoint keys.
  cases.npz            for each case: uint8 HWC input image, the reference's fp32 output
                       (runSR.sr / imageProcess.RGBFilter), its tile list (opt.iterClip()), and the
                       free-memory figure the plan was made with.
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

CKPT = {'a2': 'a2/model_new.pth', 'a3': 'a3/model_new.pth', 'a4': 'a4/model_new.pth',
        'dn_lite15': 'dn_lite15/model_new.pth', 'dn_lite5': 'dn_lite5/model_new.pth',
        'lite2': 'lite/model.pth', 'lite4': 'lite/model_4.pth', 'lite8': 'lite/model_8.pth'}


def smooth_noise_u8(h, w, seed):
  """SURVEY.md §8d recipe: bicubic-upsampled low-res noise + 0.03 randn, quantised to 8 bit (HWC)."""
  g = torch.Generator().manual_seed(seed)
  lo = torch.rand(1, 3, max(2, (h + 7) // 8), max(2, (w + 7) // 8), generator=g)
  x = torch.nn.functional.interpolate(lo, size=(h, w), mode='bicubic', align_corners=False).clamp(0, 1)[0]
  x = (x + 0.03 * torch.randn(3, h, w, generator=g)).clamp(0, 1)
  return (x * 255).round().to(torch.uint8).permute(1, 2, 0).contiguous().numpy()


CASES = [
  # name, kind, arg, (H, W), crop, seed
  ('a2_single', 'sr', 2, (64, 96), 0, 1),
  ('a2_unaligned', 'sr', 2, (45, 70), 0, 2),          # reflect-pad-to-8 on both axes
  ('a2_tiled', 'sr', 2, (72, 100), 48, 3),
  ('a3_tiled', 'sr', 3, (64, 88), 48, 4),              # pad 9
  ('a4_tiled', 'sr', 4, (64, 96), 48, 5),
  ('a4_single', 'sr', 4, (40, 56), 0, 6),
  ('dn15_tiled', 'dn', 'lite15', (72, 96), 48, 7),     # pad 7
  ('dn5_single_rgba', 'dn', 'lite5', (48, 64), 0, 8),  # alpha bypass (4th plane appended below)
  # MoeNet_lite2 (runSR.py:21-23): FRM's global mean makes the result depend on the tile, so tiled cases matter
  ('lite2_tiled', 'sr:lite', 2, (72, 100), 48, 9),
  ('lite4_single', 'sr:lite', 4, (40, 56), 0, 10),
  ('lite8_single', 'sr:lite', 8, (32, 40), 0, 11),
]


def main():
  ref = R.load()
  for name, rel in CKPT.items():
    sd = R.state_dict(rel)
    np.savez(os.path.join(HERE, 'weights_%s.npz' % name),
             **{k: v.numpy().astype(np.float16) for k, v in sd.items()})
  out, meta = {}, {}
  to_tensor = lambda im: torch.from_numpy(im).permute(2, 0, 1).float() / 255   # == torchvision to_tensor
  for name, kind, arg, (h, w), crop, seed in CASES:
    img = smooth_noise_u8(h, w, seed)
    x = to_tensor(img)
    ram = int(ref['config'].calcFreeMem())
    ref['config'].calcFreeMem = (lambda r: (lambda *a, **k: r))(ram)   # pin the plan input we record
    model = 'a'
    if kind.startswith('sr:'):
      kind, model = kind.split(':')
    if kind == 'sr':
      y, plan, opt = R.run_sr(x, arg, crop=crop, model=model)
    else:
      if name.endswith('rgba'):
        g = torch.Generator().manual_seed(seed + 100)
        x = torch.cat([x, torch.rand(1, h, w, generator=g)], 0)
      y, plan, opt = R.run_dn(x, arg, crop=crop)
    out[name + '.img'] = img
    if x.shape[0] == 4:
      out[name + '.alpha'] = x[3].numpy()
    out[name + '.ref'] = y.numpy().astype(np.float32)
    meta[name] = dict(kind=kind, arg=arg, model=model, crop=crop, ram=ram, pad=int(opt.padding), scale=int(opt.scale),
                      ram_coef=float(opt.ramCoef), tiles=[[int(v) for v in t] for t in plan])
    print(name, tuple(y.shape), len(plan), 'tiles')
  out['meta'] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
  np.savez(os.path.join(HERE, 'cases.npz'), **out)


if __name__ == '__main__':
  main()
