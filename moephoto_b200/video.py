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


class _Stage:
  """pinned host buffers + device buffers of one in-flight batch"""

  def __init__(self, batch, h, w, oh, ow, bit_depth, dev):
    dt = torch.uint8 if bit_depth <= 8 else torch.int16
    self.host_in = torch.empty((batch, h, w, 3), dtype=dt).pin_memory()
    self.host_out = torch.empty((batch, oh, ow, 3), dtype=dt).pin_memory()
    self.dev_in = torch.empty((batch, h, w, 3), dtype=dt, device=dev)
    self.dev_out = torch.empty((batch, oh, ow, 3), dtype=dt, device=dev)
    self.done = torch.cuda.Event()
    self.idx = []


def process_frames(frames, opts, bit_depth=16, swap_rb=True, batch=None, rank=0, world=1):
  """frames: sequence of HWC integer arrays (bgr48le / bgr24 as video.py pipes them when swap_rb) ; opts: the Options
  of the step chain in order (e.g. [runDN.getOpt(...), runSR.getOpt(...)]).  Yields (index, HWC integer array) for
  the frames of this rank (index % world == rank), `batch` frames per engine call.  Two batches are in flight: while
  the GPU works on one, the host stages the next into page-locked memory and hands out the previous one's frames
  (one H2D and one D2H copy per batch, the integer<->fp16 conversions run on the GPU)."""
  mine = [i for i in range(len(frames)) if i % world == rank]
  if not mine:
    return
  h, w = frames[mine[0]].shape[:2]
  total_scale = 1
  for o in opts:
    total_scale *= o.scale
  oh, ow = h * total_scale, w * total_scale
  if batch is None:
    free, _ = torch.cuda.mem_get_info()
    batch = min(16, min(max_frames_per_call(o, h * s, w * s, int(free * .7)) for o, s in zip(opts, _cum_scales(opts))))
  batch = max(1, min(batch, len(mine)))
  eng = opts[0].modelCached.engine
  dev = torch.device('cuda', eng.device_id)
  npdt = np.uint8 if bit_depth <= 8 else np.uint16
  stages = [_Stage(batch, h, w, oh, ow, bit_depth, dev) for _ in range(2)]
  stream = lambda: ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)

  def launch(st, idx):
    st.idx = idx
    hin = st.host_in.numpy().view(npdt)
    for j, i in enumerate(idx):
      hin[j] = frames[i]
    n = len(idx)
    st.dev_in[:n].copy_(st.host_in[:n], non_blocking=True)
    x = torch.empty((3 * n, h, w), dtype=torch.half, device=dev)
    for j in range(n):
      _lib.check(eng.lib.moe_to_planar_f16(eng.handle, ctypes.c_void_p(st.dev_in[j].data_ptr()), int(bit_depth), h, w, 3, int(swap_rb),
                                           ctypes.c_void_p(x[3 * j].data_ptr()), stream()))
    for opt in opts:
      x = run_step(opt, x)
    for j in range(n):
      _lib.check(eng.lib.moe_to_output(eng.handle, ctypes.c_void_p(x[3 * j].data_ptr()), int(bit_depth), oh, ow, 3, int(swap_rb),
                                       ctypes.c_void_p(st.dev_out[j].data_ptr()), stream()))
    st.host_out[:n].copy_(st.dev_out[:n], non_blocking=True)
    st.done.record(torch.cuda.current_stream(dev))

  def collect(st):
    st.done.synchronize()
    hout = st.host_out.numpy().view(npdt)
    for j, i in enumerate(st.idx):
      yield i, hout[j].copy()

  chunks = [mine[k:k + batch] for k in range(0, len(mine), batch)]
  launch(stages[0], chunks[0])
  for k in range(len(chunks)):
    if k + 1 < len(chunks):
      launch(stages[(k + 1) % 2], chunks[k + 1])
    yield from collect(stages[k % 2])


def _cum_scales(opts):
  s, out = 1, []
  for o in opts:
    out.append(s)
    s *= o.scale
  return out
