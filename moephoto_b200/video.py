"""Frame-batched driver for the video route (reference loop: video.py:323-397).

MoePhoto feeds its SR / DN closures ONE frame per Python iteration (`process((raw,h,w))`, video.py:351-360).  The
networks treat planes as a batch, so B frames can go through the engine as 3B planes in one call — with the tile plan
of a SINGLE frame, which keeps every frame's result bit-identical to the per-frame call (the reference's own batched
plan would differ: its pixel budget shrinks with the plane count squared, imageProcess.py:74-75).  Frames are
independent, so on several GPUs each rank simply takes every world-th frame: no collective on the data path.
"""
import ctypes

import numpy as np
import torch

from . import imageProcess as IP
from . import _lib


def _plan_for_one_frame(opt, h, w):
  if opt.iterClip is None or opt.plan is None or (opt.plan.in_h, opt.plan.in_w) != (h, w) or opt.outShape[0] != 3:
    opt.iterClip = None
    IP.prepareOpt(opt, (3, h, w))
  return opt.plan


def max_frames_per_call(opt, h, w, free_bytes):
  """how many frames fit the workspace at once (the engine reports its own need)"""
  plan = _plan_for_one_frame(opt, h, w)
  eng = opt.modelCached.engine
  per_frame = eng.lib.moe_plan_workspace_bytes(opt.modelCached.handle, 3, ctypes.byref(plan.c), 0, plan.out_h)
  return max(1, int(free_bytes // max(1, per_frame)))


def run_step(opt, x, strength=None):
  """one SR or DN step on a plane batch x (3B,H,W) with the single-frame tile plan; strengthOp applied for DN"""
  plan = _plan_for_one_frame(opt, x.shape[-2], x.shape[-1])
  y = IP.run_plan(opt.modelCached, x, plan)
  s = opt.strength if strength is None else strength
  return IP.strengthOp(y, x, s) if (plan.scale == 1 and s != 1) else y


def process_frames(frames, opts, bit_depth=16, swap_rb=True, batch=None, rank=0, world=1):
  """frames: sequence of HWC integer arrays (bgr48le / bgr24 as video.py pipes them when swap_rb) ; opts: the Options
  of the step chain in order (e.g. [runDN.getOpt(...), runSR.getOpt(...)]).  Yields (index, HWC integer array) for
  the frames of this rank (index % world == rank), processed `batch` frames per engine call."""
  mine = [i for i in range(len(frames)) if i % world == rank]
  if not mine:
    return
  h, w = frames[mine[0]].shape[:2]
  if batch is None:
    free, _ = torch.cuda.mem_get_info()
    batch = min(16, min(max_frames_per_call(o, h * s, w * s, int(free * .8)) for o, s in zip(opts, _cum_scales(opts))))
  load, store = IP.toTorch(bit_depth, swapRB=swap_rb), IP.toOutput(bit_depth, swapRB=swap_rb)
  for k in range(0, len(mine), batch):
    idx = mine[k:k + batch]
    x = torch.cat([load(frames[i]) for i in idx], 0)
    for opt in opts:
      x = run_step(opt, x)
    for j, i in enumerate(idx):
      yield i, store(x[3 * j:3 * j + 3])


def _cum_scales(opts):
  s, out = 1, []
  for o in opts:
    out.append(s)
    s *= o.scale
  return out
