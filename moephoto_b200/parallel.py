"""Row-band sharding of ONE image over the GPUs of a box (new work: the reference is single-GPU,
SURVEY.md §2c / §8e).

Every rank owns a horizontal band of the output canvas.  It runs the SAME reference tile plan, but each
tile only on the band's rows plus a 16-LR-pixel recompute halo (moe_run_plan's row window) — exact,
because the network's receptive-field radius is 15.75 LR px and every blend reads only canvas rows
inside the band.  The data path has two exchange steps and nothing else:
  1. broadcast of the LR frame from the root (4K fp16: 50 MB — cheaper than scattering haloed bands);
  2. gather of the upscaled bands into the root's canvas (grouped NCCL send/recv, one message per plane
     and rank).
One process per GPU, torch.distributed for the plumbing (backend nccl; gloo on CPU in the tests).
"""
import torch
import torch.distributed as dist


def band_rows(in_h, scale, world, rank):
  """canvas rows [lo,hi) of `rank`: LR rows are split as evenly as integers allow, then scaled"""
  lo = in_h * rank // world
  hi = in_h * (rank + 1) // world
  return lo * scale, hi * scale


def agree_on_free_memory(local_free, device, group=None):
  """the tile plan depends on free memory (imageProcess.py:136-138): every rank must use the same
  figure or the plans — and so the seams — differ.  min over ranks."""
  t = torch.tensor([float(local_free)], dtype=torch.float64, device=device)
  dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)
  return int(t.item())


def sharded_run(compute_band, x, planes, in_h, in_w, scale, root=0, group=None, out=None, gather=True):
  """x: (planes,in_h,in_w) tensor, valid on `root` (other ranks pass a same-shaped buffer to receive
  into).  compute_band(x, lo, hi, canvas) must fill canvas[:, lo:hi, :].  Returns the full canvas on
  the root, the local band's canvas elsewhere."""
  world = dist.get_world_size(group)
  rank = dist.get_rank(group)
  if world > 1:
    dist.broadcast(x, src=dist.get_global_rank(group, root) if group is not None else root, group=group)
  if out is None:
    out = x.new_empty((planes, in_h * scale, in_w * scale))
  lo, hi = band_rows(in_h, scale, world, rank)
  if hi > lo:
    compute_band(x, lo, hi, out)
  if world == 1 or not gather:
    return out
  ops = []
  g = lambda r: dist.get_global_rank(group, r) if group is not None else r
  if rank == root:
    for r in range(world):
      if r == root:
        continue
      rlo, rhi = band_rows(in_h, scale, world, r)
      if rhi > rlo:
        ops += [dist.P2POp(dist.irecv, out[p, rlo:rhi], g(r), group) for p in range(planes)]
  elif hi > lo:
    ops += [dist.P2POp(dist.isend, out[p, lo:hi], g(root), group) for p in range(planes)]
  if ops:
    for req in dist.batch_isend_irecv(ops):
      req.wait()
  return out


def sharded_doCrop(opt, x, root=0, group=None, gather=True):
  """doCrop (imageProcess.py:157-172) with the canvas rows sharded over the process group.
  All ranks call it with the same `opt` settings; `x` holds the image on the root."""
  from . import imageProcess as IP
  from .config import config
  if opt.iterClip is None or opt.count > 28 or x.shape[0] != opt.outShape[0]:
    if config.freeMemOverride is None and dist.get_world_size(group) > 1:
      config.freeMemOverride = agree_on_free_memory(config.calcFreeMem(), x.device, group)
      try:
        IP.prepareOpt(opt, x.shape)
      finally:
        config.freeMemOverride = None
    else:
      IP.prepareOpt(opt, x.shape)
  else:
    opt.count += 1
  plan = opt.plan
  opt.outShape[0] = x.size(0)
  run = lambda xi, lo, hi, canvas: IP.run_plan(opt.modelCached, xi, plan, canvas, rows=(lo, hi))
  return sharded_run(run, x, x.shape[0], plan.in_h, plan.in_w, plan.scale, root, group, None, gather).detach()


class SharedHostFrame:
  """A host frame every rank of the box can write: a /dev/shm mapping (MoePhoto hands frames between its processes
  the same way, server.py:369-372 / MoePhoto.py:10-17), page-locked in every process so each GPU copies ITS band of
  the result over ITS OWN PCIe link instead of funnelling the whole frame through rank 0."""

  def __init__(self, name, shape, dtype=torch.uint8, create=False):
    import numpy as np
    self.path = '/dev/shm/' + name
    nbytes = int(np.prod(shape)) * torch.empty((), dtype=dtype).element_size()
    if create:
      with open(self.path, 'wb') as f:
        f.truncate(nbytes)
    self._map = np.memmap(self.path, dtype=np.uint8, mode='r+', shape=(nbytes,))
    self.tensor = torch.from_numpy(self._map).view(dtype).view(*shape)
    self._registered = False
    if torch.cuda.is_available():
      rc = torch.cuda.cudart().cudaHostRegister(self.tensor.data_ptr(), nbytes, 0)
      self._registered = int(rc) == 0

  def close(self, unlink=False):
    if self._registered:
      torch.cuda.cudart().cudaHostUnregister(self.tensor.data_ptr())
      self._registered = False
    if unlink:
      import os
      try:
        os.unlink(self.path)
      except OSError:
        pass


def sharded_enhance_host(opt, host_in, shared_out, bits_in=8, bits_out=8, root=0, group=None):
  """host uint8/uint16 HWC frame on the root -> HWC result in `shared_out` (a SharedHostFrame every rank opened):
  toTorch on the root, broadcast, every rank computes its row band and converts + copies it to the host itself.
  No gather: the assembled frame only ever exists in host memory."""
  import ctypes
  from . import imageProcess as IP, _lib
  rank = dist.get_rank(group)
  dev = torch.device('cuda', torch.cuda.current_device())
  shape = [host_in.shape[0], host_in.shape[1]] if rank == root else [0, 0]
  if opt.plan is None or rank == root and (opt.plan.in_h, opt.plan.in_w) != tuple(shape):
    pass                                                           # the root's frame decides; tell the others below
  box = [shape]
  dist.broadcast_object_list(box, src=dist.get_global_rank(group, root) if group is not None else root, group=group)
  h, w = box[0]
  x = IP.toTorch(bits_in)(host_in) if rank == root else torch.empty((3, h, w), dtype=torch.half, device=dev)
  y = sharded_doCrop(opt, x, root, group, gather=False)
  plan = opt.plan
  lo, hi = band_rows(plan.in_h, plan.scale, dist.get_world_size(group), rank)
  if hi > lo:
    band = y[:, lo:hi].contiguous()
    eng = opt.modelCached.engine
    q = torch.empty((hi - lo, plan.out_w, 3), dtype=torch.uint8 if bits_out <= 8 else torch.int16, device=dev)
    _lib.check(eng.lib.moe_to_output(eng.handle, ctypes.c_void_p(band.data_ptr()), int(bits_out), hi - lo, plan.out_w, 3, 0,
                                     ctypes.c_void_p(q.data_ptr()), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
    shared_out.tensor[lo:hi].copy_(q.view(shared_out.tensor.dtype) if q.dtype != shared_out.tensor.dtype else q, non_blocking=True)
  torch.cuda.current_stream().synchronize()
  dist.barrier(group)                     # every band has landed in host memory
