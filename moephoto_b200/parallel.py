"""Row-band sharding of ONE image over the GPUs of a box (new work: the reference is single-GPU,
SURVEY.md §2c / §8e).

Every rank owns a horizontal band of the output canvas.  It runs the SAME reference tile plan, but each
tile only on the band's rows plus a 16-LR-pixel recompute halo (moe_run_plan's row window) — exact,
because the network's receptive-field radius is 15.75 LR px and every blend reads only canvas rows
inside the band.  The data path has two exchange steps and nothing else: the LR frame must reach every rank and
the upscaled bands must reach the root.
  * BandSharder (the default on one NVLink box): both exchanges are fused into the path's own kernels over peer memory —
    every rank reads the frame from, and stores its band into, the ROOT's HBM through IPC-mapped pointers;
  * sharded_run / sharded_doCrop (the portable form, also the CPU test vehicle with gloo): NCCL broadcast of the frame
    (4K fp16: 50 MB), grouped NCCL send/recv gather of the bands (one message per plane and rank).
One process per GPU, torch.distributed for the plumbing (backend nccl; gloo on CPU in the tests).
"""
import ctypes

import torch
import torch.distributed as dist


HALO_LR = 16   # recompute halo of a band in LR rows = the receptive-field radius of the nets (csrc/engine.cu kHaloLR)


def band_rows(in_h, scale, world, rank):
  """canvas rows [lo,hi) of `rank`.  What a rank COMPUTES is its band plus a 16-row halo on every side that is not an image
  edge, so the LR rows are split such that the computed row counts are equal (the two edge bands keep 16 rows more than the
  interior ones); images too small for that fall back to the even split.  Any partition gives the same bits."""
  if world > 1:
    total = in_h + HALO_LR * (2 * world - 2)
    edges = [0]
    for r in range(world):
      halos = HALO_LR * ((r > 0) + (r < world - 1))
      edges.append(edges[-1] + total * (r + 1) // world - total * r // world - halos)
    if edges[-1] == in_h and all(b - a >= 1 for a, b in zip(edges, edges[1:])):
      return edges[rank] * scale, edges[rank + 1] * scale
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
  _agreed_plan(opt, x.shape, x.device, group)
  plan = opt.plan
  opt.outShape[0] = x.size(0)
  run = lambda xi, lo, hi, canvas: IP.run_plan(opt.modelCached, xi, plan, canvas, rows=(lo, hi))
  return sharded_run(run, x, x.shape[0], plan.in_h, plan.in_w, plan.scale, root, group, None, gather).detach()


class _Cudart:
  """the few CUDA runtime calls torch does not expose, through ctypes: raw allocations whose IPC handle can be opened by
  another process IN THE CONTEXT OF ITS OWN GPU (torch's own tensor sharing opens the handle in the exporting device's
  context: the importer's copy engines can then reach the memory, its SMs cannot)"""

  class Handle(ctypes.Structure):
    _fields_ = [('reserved', ctypes.c_char * 64)]

  def __init__(self):
    rt = None
    for name in ('libcudart.so.12', 'libcudart.so'):
      try:
        rt = ctypes.CDLL(name)
        break
      except OSError:
        continue
    if rt is None:
      raise RuntimeError('libcudart not found')
    rt.cudaMalloc.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_size_t]
    rt.cudaFree.argtypes = [ctypes.c_void_p]
    rt.cudaIpcGetMemHandle.argtypes = [ctypes.POINTER(self.Handle), ctypes.c_void_p]
    rt.cudaIpcOpenMemHandle.argtypes = [ctypes.POINTER(ctypes.c_void_p), self.Handle, ctypes.c_uint]
    rt.cudaIpcCloseMemHandle.argtypes = [ctypes.c_void_p]
    self.rt = rt

  def _check(self, rc, what):
    if rc != 0:
      self.rt.cudaGetLastError()
      raise RuntimeError('%s failed with CUDA error %d' % (what, rc))

  def malloc(self, nbytes):
    p = ctypes.c_void_p()
    self._check(self.rt.cudaMalloc(ctypes.byref(p), nbytes), 'cudaMalloc(%d)' % nbytes)
    return p.value

  def export(self, ptr):
    h = self.Handle()
    self._check(self.rt.cudaIpcGetMemHandle(ctypes.byref(h), ptr), 'cudaIpcGetMemHandle')
    return ctypes.string_at(ctypes.byref(h), 64)

  def open(self, raw):
    h = self.Handle()
    ctypes.memmove(ctypes.byref(h), raw, 64)
    p = ctypes.c_void_p()
    self._check(self.rt.cudaIpcOpenMemHandle(ctypes.byref(p), h, 1), 'cudaIpcOpenMemHandle')   # 1 = cudaIpcMemLazyEnablePeerAccess
    return p.value


class _RawCudaArray:
  """a raw device pointer as something torch.as_tensor can wrap without copying"""

  def __init__(self, ptr, shape, typestr='<f2'):
    self.__cuda_array_interface__ = {'shape': tuple(shape), 'typestr': typestr, 'data': (int(ptr), False), 'version': 2}


def _agreed_plan(opt, shape, device, group=None):
  """prepareOpt on every rank with ONE free-memory figure (the minimum over the ranks): the tile plan depends on it
  (imageProcess.py:136-138) and every rank must cut the same tiles.  A rank that cannot probe its memory makes all raise."""
  from . import imageProcess as IP
  from .config import config
  if not IP._plan_is_stale(opt, shape):
    opt.count += 1
    return
  if config.freeMemOverride is None and dist.get_world_size(group) > 1:
    try:
      local_free = float(config.calcFreeMem())
    except Exception:
      local_free = -1.0
    t = torch.tensor([local_free], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)
    if t.item() < 0:
      raise MemoryError('Can not calculate free memory.')
    config.freeMemOverride = int(t.item())
    try:
      IP.prepareOpt(opt, shape)
    finally:
      config.freeMemOverride = None
  else:
    IP.prepareOpt(opt, shape)


class BandSharder:
  """One image over the GPUs of a box with NO bulk collective: the root's input frame and output canvas are mapped into every
  rank's address space (CUDA IPC over NVLink / NVSwitch peer memory), and each rank's kernels read and write them in place —
  conv_first_kernel pulls the band's LR rows (+ halo) straight from the root's frame, head_stencil_kernel stores (and, on the
  seams, blends) the band straight into the root's canvas.  The transfers ARE the first and last kernel of every tile, so they
  overlap the tensor-core work of the neighbouring tiles; what remains on the stream per step is two one-element collectives that
  order the ranks (frame ready -> go, bands stored -> done).  Round 1 broadcast the frame and gathered the bands with NCCL after
  the compute: 2.9 of 14.1 ms per frame at 8 GPUs.
  Every rank builds one with the same `opt` settings and calls run() / run_host() collectively."""

  def __init__(self, opt, shape, device, root=0, group=None):
    from . import imageProcess as IP
    self.opt, self.root, self.group = opt, root, group
    self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
    self.device = device
    _agreed_plan(opt, shape, device, group)
    plan = self.plan = opt.plan
    opt.outShape[0] = shape[0]
    self.planes = shape[0]
    self.lo, self.hi = band_rows(plan.in_h, plan.scale, self.world, self.rank)
    self.token = torch.zeros(1, dtype=torch.int32, device=device)
    self.peer = False
    self.x = self.canvas = None
    src = dist.get_global_rank(group, root) if group is not None else root
    box = [None]
    xs, cs = (self.planes, plan.in_h, plan.in_w), (self.planes, plan.out_h, plan.out_w)
    nb = lambda shp: 2 * shp[0] * shp[1] * shp[2]
    self._raw = None
    try:
      with torch.cuda.device(device):
        torch.cuda.current_stream()                                            # the primary context of THIS GPU is current
        rt = _Cudart()
        if self.rank == root:
          self._raw = (rt, rt.malloc(nb(xs)), rt.malloc(nb(cs)))
          box = [(rt.export(self._raw[1]), rt.export(self._raw[2]))]
    except Exception as ex:
      import sys
      print('moephoto_b200: cannot export the root buffers over CUDA IPC (%r)' % (ex,), file=sys.stderr, flush=True)
      box = [None]
    dist.broadcast_object_list(box, src=src, group=group)
    ok = box[0] is not None
    if ok:
      try:
        with torch.cuda.device(device):
          px, pc = (self._raw[1], self._raw[2]) if self.rank == root else (rt.open(box[0][0]), rt.open(box[0][1]))
          self.x = torch.as_tensor(_RawCudaArray(px, xs), device=device)
          self.canvas = torch.as_tensor(_RawCudaArray(pc, cs), device=device)
          if self.x.data_ptr() != px or self.canvas.data_ptr() != pc:
            raise RuntimeError('torch copied the raw buffer instead of wrapping it')
          probe = self.canvas.view(-1)[:8].clone()                             # fails loudly here, not inside a convolution
          torch.cuda.synchronize(device)
      except Exception as ex:
        import sys
        print('moephoto_b200: peer mapping of the root buffers failed on rank %d (%r); falling back to NCCL broadcast + gather' % (self.rank, ex),
              file=sys.stderr, flush=True)
        ok = False
    flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=device)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
    self.peer = bool(flag.item())
    if not self.peer:                                                          # fall-back: NCCL broadcast + gather (sharded_run)
      self.x = torch.empty(xs, dtype=torch.half, device=device)
      self.canvas = None
    self.band = None                                                           # local canvas of the host-to-host path

  def _order(self, op):
    """a one-element collective on the current stream: orders every rank's stream after the root's (broadcast) or the
    root's after every rank's (all_reduce)"""
    if op == 'go':
      dist.broadcast(self.token, src=dist.get_global_rank(self.group, self.root) if self.group is not None else self.root, group=self.group)
    else:
      dist.all_reduce(self.token, group=self.group)

  def run(self, x=None):
    """x: the (planes,H,W) frame on the root (None: the caller already wrote it into self.x).  Returns the stitched canvas
    on the root (a buffer owned by the sharder, overwritten by the next call), None elsewhere."""
    from . import imageProcess as IP
    if not self.peer:
      xin = x if self.rank == self.root else self.x
      if self.rank == self.root and x is None:
        xin = self.x
      return sharded_doCrop(self.opt, xin, self.root, self.group)
    if self.rank == self.root and x is not None and x.data_ptr() != self.x.data_ptr():
      self.x.copy_(x)
    self._order('go')                                                          # the frame is complete in the root's memory
    if self.hi > self.lo:
      IP.run_plan(self.opt.modelCached, self.x, self.plan, self.canvas, rows=(self.lo, self.hi))
    self._order('done')                                                        # every band has been stored into the root's canvas
    return self.canvas if self.rank == self.root else None

  def run_host(self, host_in, shared_out, bits_in=8, bits_out=8):
    """host integer HWC frame on the root -> HWC result in `shared_out` (a SharedHostFrame every rank opened).  The root
    uploads and converts the frame once; every rank pulls its band's rows from the root's memory, computes the band into a
    LOCAL canvas, converts it and copies it to the host over its own PCIe link."""
    import ctypes
    from . import imageProcess as IP, _lib
    if not self.peer:
      return sharded_enhance_host(self.opt, host_in, shared_out, bits_in, bits_out, self.root, self.group)
    eng, plan = self.opt.modelCached.engine, self.plan
    stream = lambda: ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
    if self.rank == self.root:
      raw = torch.from_numpy(host_in).to(self.device, non_blocking=True)
      _lib.check(eng.lib.moe_to_planar_f16(eng.handle, ctypes.c_void_p(raw.data_ptr()), int(bits_in), plan.in_h, plan.in_w, self.planes, 0,
                                           ctypes.c_void_p(self.x.data_ptr()), stream()))
    self._order('go')
    if self.hi > self.lo:
      # compute the band, convert and copy each tile's finished columns to the host under the next tile's compute
      _lib.check(eng.lib.moe_run_band_to_host(self.opt.modelCached.handle, ctypes.c_void_p(self.x.data_ptr()), self.x.stride(0), self.x.stride(1),
                                              self.planes, ctypes.byref(plan.c), self.lo, self.hi, ctypes.c_void_p(shared_out.tensor.data_ptr()),
                                              int(bits_out), stream()))
    torch.cuda.current_stream(self.device).synchronize()
    dist.barrier(self.group)                # every band has landed in host memory


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
  box = [shape]                                                    # the root's frame decides the size; tell the others
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
