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


class FrameBatcher:
  """Keeps the page-locked staging buffers (allocating ~0.5 GB of pinned memory costs more than processing a batch)
  and the device buffers alive across calls; `process(frames)` is the frame loop of video.py:349-360."""

  def __init__(self, opts, height, width, bit_depth=16, swap_rb=True, batch=None):
    self.opts, self.h, self.w, self.bit_depth, self.swap_rb = list(opts), height, width, bit_depth, swap_rb
    scale = 1
    for o in self.opts:
      scale *= o.scale
    self.oh, self.ow = height * scale, width * scale
    if batch is None:
      free, _ = torch.cuda.mem_get_info()
      batch = min(16, min(max_frames_per_call(o, height * s, width * s, int(free * .7)) for o, s in zip(self.opts, _cum_scales(self.opts))))
    self.batch = max(1, int(batch))
    self.eng = self.opts[0].modelCached.engine
    self.dev = torch.device('cuda', self.eng.device_id)
    self.stages = [_Stage(self.batch, height, width, self.oh, self.ow, bit_depth, self.dev) for _ in range(2)]

  def _launch(self, st, frames, idx):
    eng, dev, h, w = self.eng, self.dev, self.h, self.w
    stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    npdt = np.uint8 if self.bit_depth <= 8 else np.uint16
    st.idx = idx
    hin = st.host_in.numpy().view(npdt)
    for j, i in enumerate(idx):
      hin[j] = frames[i]
    n = len(idx)
    st.dev_in[:n].copy_(st.host_in[:n], non_blocking=True)
    x = torch.empty((3 * n, h, w), dtype=torch.half, device=dev)
    for j in range(n):
      _lib.check(eng.lib.moe_to_planar_f16(eng.handle, ctypes.c_void_p(st.dev_in[j].data_ptr()), int(self.bit_depth), h, w, 3, int(self.swap_rb),
                                           ctypes.c_void_p(x[3 * j].data_ptr()), stream))
    for opt in self.opts:
      x = run_step(opt, x)
    for j in range(n):
      _lib.check(eng.lib.moe_to_output(eng.handle, ctypes.c_void_p(x[3 * j].data_ptr()), int(self.bit_depth), self.oh, self.ow, 3, int(self.swap_rb),
                                       ctypes.c_void_p(st.dev_out[j].data_ptr()), stream))
    st.host_out[:n].copy_(st.dev_out[:n], non_blocking=True)
    st.done.record(torch.cuda.current_stream(dev))

  def _collect(self, st, copy):
    st.done.synchronize()
    hout = st.host_out.numpy().view(np.uint8 if self.bit_depth <= 8 else np.uint16)
    for j, i in enumerate(st.idx):
      yield i, (hout[j].copy() if copy else hout[j])

  def process(self, frames, rank=0, world=1, copy=True):
    """yields (index, HWC integer frame) for the frames with index % world == rank.  Two batches are in flight: while
    the GPU works on one, the host stages the next into page-locked memory and hands out the previous one's frames.
    copy=False yields views of the staging buffer, valid until the batch after next is launched."""
    mine = [i for i in range(len(frames)) if i % world == rank]
    chunks = [mine[k:k + self.batch] for k in range(0, len(mine), self.batch)]
    if not chunks:
      return
    self._launch(self.stages[0], frames, chunks[0])
    for k in range(len(chunks)):
      if k + 1 < len(chunks):
        self._launch(self.stages[(k + 1) % 2], frames, chunks[k + 1])
      yield from self._collect(self.stages[k % 2], copy)


class _PipeFrames:
  """the frames of a raw byte pipe as a lazily growing sequence (what FrameBatcher.process indexes)"""

  def __init__(self, read, height, width, bit_depth, limit=None):
    self.read, self.h, self.w, self.limit = read, height, width, limit
    self.dtype = np.uint8 if bit_depth <= 8 else np.uint16
    self.nbytes = height * width * 3 * (1 if bit_depth <= 8 else 2)
    self.frames, self.eof = [], False

  def fill(self, n):
    """read until n frames are buffered or the pipe ends (an empty or short read: video.py:351-353)"""
    while not self.eof and len(self.frames) < n and (self.limit is None or len(self.frames) < self.limit):
      raw = self.read(self.nbytes)
      if raw is None or len(raw) < self.nbytes:
        self.eof = True
        break
      self.frames.append(np.frombuffer(raw, dtype=self.dtype).reshape(self.h, self.w, 3))
    return len(self.frames)


def pipe_loop(read, write, height, width, opts, bit_depth=16, swap_rb=True, batch=None, stop=None, should_stop=None):
  """The reference's frame loop (video.py:339-360: `raw_image = procIn.stdout.read(frameBytes)` -> process((raw, h, w)) ->
  `procOut.stdin.write(buffer)`) with B frames per engine call instead of one.  `read(nbytes)` and `write(bytes)` are the two pipe
  ends (ffmpeg's stdout / stdin in MoePhoto; any callables here); frames are bgr48le (bit_depth 16) or bgr24 as MoePhoto's ffmpeg
  command lines produce them, `opts` the Options of the step chain (toNumPy -> toTorch -> DN / SR ... -> toOutput -> BGR -> toBuffer
  collapse into the batcher's two conversion kernels).  Frames are written in order; returns the number of frames processed.
  `should_stop()` is polled once per batch (context.stopFlag, video.py:350)."""
  src = _PipeFrames(read, height, width, bit_depth, stop)
  if src.fill(1) == 0:
    return 0
  fb = FrameBatcher(opts, height, width, bit_depth, swap_rb, batch)
  done = 0
  while True:
    n = src.fill(done + fb.batch)
    if n == done or (should_stop is not None and should_stop()):
      break
    for _, frame in fb.process(src.frames[done:n], copy=False):
      write(frame.tobytes())
    for i in range(done, n):
      src.frames[i] = None                                                     # drop the input bytes of finished frames
    done = n
  return done


def process_frames(frames, opts, bit_depth=16, swap_rb=True, batch=None, rank=0, world=1):
  """one-shot convenience around FrameBatcher: frames = sequence of HWC integer arrays (bgr48le / bgr24 as video.py
  pipes them when swap_rb); opts = the Options of the step chain in order; yields (index, HWC integer frame)"""
  if len(frames) == 0:
    return
  h, w = frames[0].shape[:2]
  mine = len([i for i in range(len(frames)) if i % world == rank])
  fb = FrameBatcher(opts, h, w, bit_depth, swap_rb, batch if batch is None else min(batch, max(1, mine)))
  yield from fb.process(frames, rank, world)


def _cum_scales(opts):
  s, out = 1, []
  for o in opts:
    out.append(s)
    s *= o.scale
  return out
