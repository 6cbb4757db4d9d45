"""bench.py — the BASELINE.json metric on the BASELINE.json config.

  metric : output MPix/s of 4x SR (a4 = models.Net4x) on one 3840x2160 RGB frame -> 15360x8640
  step   : one frame through runSR.sr(getOpt({'model':'a','scale':4})) — the reference's auto tile plan
           for a 180 GB GPU (4 column strips of 2160x968, pad 5, seam 20 px), every tile through the 20
           convolutions of Net4x, seam-blended and stitched on the device.
  N GPUs : the SAME frame, canvas rows sharded over the ranks (moephoto_b200/parallel.py): broadcast of
           the LR frame, per-rank row band + 16-px recompute halo, NCCL gather into rank 0 -> "strong".
  value  : frame resident in HBM (fp16 planar) -> stitched fp16 canvas resident in HBM on rank 0.
  e2e    : uint8 HWC frame in pinned HOST memory -> uint8 HWC result in HOST memory, through the C-ABI
           (moe_enhance_host at N=1; at N>1 parallel.sharded_enhance_host: every rank converts and copies its own
           band into a shared page-locked host frame), copies timed.
  --impl reference : the reference's CPU path (PyTorch conv2d on the host cores, fp32) restated by the
           oracle port (oracle/net.py forward_torch + oracle/tiling.py), all host threads, on a
           bounded sample of the same workload (a 256x256 crop per step).

Prints ONE JSON line on rank 0.  Synthetic data, real a4 weights when tests/golden/weights_a4.npz is
present (it is committed), else seeded random weights of the same architecture.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, 'tests')):
  if p not in sys.path:
    sys.path.insert(0, p)

METRIC = 'output MPix/s, 4x SR (a4) on 3840x2160 RGB'
H_IN, W_IN, SCALE = 2160, 3840, 4
FLOP_PER_LR_PIXEL_PLANE = 3945600
CPU_SAMPLE = int(os.environ.get('MOE_BENCH_CPU_SAMPLE', '256'))   # the CPU legs run a CPU_SAMPLE^2 crop per step


def a4_weights():
  path = os.path.join(ROOT, 'tests', 'golden', 'weights_a4.npz')
  if os.path.exists(path):
    return {k: v.astype(np.float32) for k, v in np.load(path).items()}, 'reference checkpoint model/a4 (fp16 copy)'
  rng = np.random.default_rng(0)
  sd = {'conv_input.weight': rng.normal(0, .05, (64, 1, 3, 3)), 'conv_input2.weight': rng.normal(0, .03, (64, 64, 3, 3)),
        'relu.weight': np.array([.1])}
  for i in range(1, 7):
    p = 'convt_F%d.0.' % i
    sd.update({p + 'conv_1.weight': rng.normal(0, .03, (64, 64, 3, 3)), p + 'relu.weight': np.array([.1]),
               p + 'conv_2.weight': rng.normal(0, .03, (64, 64, 3, 3)), p + 'scale.scale': np.array([.25])})
  for b in ('u', 'convt_R1'):
    for s in range(2):
      sd.update({'%s.%d.0.weight' % (b, s): rng.normal(0, .03, (256, 64, 3, 3)), '%s.%d.0.bias' % (b, s): np.zeros(256),
                 '%s.%d.2.weight' % (b, s): np.array([.1])})
    sd['%s.2.weight' % b] = rng.normal(0, .03, (1, 64, 3, 3))
  return {k: np.asarray(v, dtype=np.float32) for k, v in sd.items()}, 'random init'


def synthetic_frame(h, w, seed=0):
  """uint8 HWC: smooth low-frequency content + noise (SURVEY.md §8d recipe, numpy only)"""
  rng = np.random.default_rng(seed)
  lo = rng.random((h // 8 + 2, w // 8 + 2, 3)).astype(np.float32)
  yy = np.linspace(0, lo.shape[0] - 1.001, h)
  xx = np.linspace(0, lo.shape[1] - 1.001, w)
  y0, x0 = yy.astype(int), xx.astype(int)
  fy, fx = (yy - y0)[:, None, None], (xx - x0)[None, :, None]
  img = (lo[y0][:, x0] * (1 - fy) * (1 - fx) + lo[y0 + 1][:, x0] * fy * (1 - fx) +
         lo[y0][:, x0 + 1] * (1 - fy) * fx + lo[y0 + 1][:, x0 + 1] * fy * fx)
  img = np.clip(img + 0.03 * rng.standard_normal(img.shape).astype(np.float32), 0, 1)
  return np.ascontiguousarray((img * 255).round().astype(np.uint8))


def host_threads():
  """cores this process may really use: affinity mask, capped by the cgroup CPU quota"""
  n = len(os.sched_getaffinity(0)) if hasattr(os, 'sched_getaffinity') else (os.cpu_count() or 1)
  try:
    q, per = open('/sys/fs/cgroup/cpu.max').read().split()
    if q != 'max':
      n = max(1, min(n, int(float(q) / float(per))))
  except Exception:
    try:
      q = int(open('/sys/fs/cgroup/cpu/cpu.cfs_quota_us').read())
      per = int(open('/sys/fs/cgroup/cpu/cpu.cfs_period_us').read())
      if q > 0:
        n = max(1, min(n, q // per))
    except Exception:
      pass
  return n


def peaks():
  path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
  if os.path.exists(path):
    p = json.load(open(path))
    return dict(tflops=p.get('bf16_tflops_sustained', p.get('bf16_tflops')), hbm=p.get('hbm_gbs'), src='measured (MEASURED_PEAKS.json, sustained cuBLAS bf16)')
  return dict(tflops=1400.0, hbm=6650.0, src='fallback (B200_PROFILING.md)')


class ClockSampler:
  """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)"""
  Q = 'index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,' \
      'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

  def __init__(self, gpu_index):
    self.f = tempfile.NamedTemporaryFile('w+', suffix='.csv', delete=False)
    self.p = None
    try:
      self.p = subprocess.Popen(['nvidia-smi', '-i', str(gpu_index), '--query-gpu=' + self.Q, '--format=csv,noheader,nounits', '-lms', '50'],
                                stdout=self.f, stderr=subprocess.DEVNULL)
    except Exception:
      pass

  def mark(self):
    """samples written before this call belong to the warm-up and are dropped"""
    self.f.flush()
    try:
      self.skip = sum(1 for _ in open(self.f.name))
    except Exception:
      self.skip = 0

  def stop(self):
    out = dict(sm_mhz=None, sm_max_mhz=None, reasons=[], samples=0)
    if self.p is None:
      return out
    self.p.terminate()
    try:
      self.p.wait(timeout=5)
    except Exception:
      self.p.kill()
    self.f.flush()
    rows = [r.strip().split(', ') for r in open(self.f.name) if r.strip()][getattr(self, 'skip', 0):]
    os.unlink(self.f.name)
    sm, reasons = [], set()
    for r in rows:
      try:
        sm.append(float(r[1])); out['sm_max_mhz'] = float(r[2])
      except Exception:
        continue
      for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), r[5:9]):
        if v.strip().lower().startswith('active'):
          reasons.add(name)
    if sm:
      sm.sort()
      busy = [v for v in sm if v > 0.5 * sm[-1]] or sm
      out['sm_mhz'] = busy[len(busy) // 2]
    out['reasons'], out['samples'] = sorted(reasons), len(sm)
    return out


# ------------------------------------------------------------------------------------------------
def cpu_port_step(sd, crop_u8, threads):
  """one bounded sample of the workload on the host cores: a4 over a CPU_SAMPLE^2 crop, fp32, through
  the oracle's restatement of doCrop + MyNet.forward with PyTorch's CPU conv2d (what the reference's
  CPU path executes).  Returns seconds."""
  from oracle import net as ONet, tiling as OT
  x = OT.to_planar(crop_u8, 8)
  plan = OT.make_plan(x.shape, 8e9, .9 / 41951.3, 5, SCALE, 8, 0)     # ramCoef of the CPU fp32 row, runSR.py:9
  t0 = time.perf_counter()
  y = OT.do_crop(lambda a: ONet.forward_torch(sd, a), x, plan)
  OT.to_output(y, 8)
  return time.perf_counter() - t0


def run_reference(args, rank):
  if rank != 0:
    return
  import torch
  threads = host_threads()
  torch.set_num_threads(threads)
  sd, wsrc = a4_weights()
  crop = synthetic_frame(CPU_SAMPLE, CPU_SAMPLE, 1)
  for _ in range(max(1, args.warmup)):
    cpu_port_step(sd, crop, threads)
  t = [cpu_port_step(sd, crop, threads) for _ in range(args.steps)]
  sec = sum(t) / len(t)
  mpix = (CPU_SAMPLE * SCALE) ** 2 / 1e6
  val = mpix / sec
  sample = '%dx%d RGB crop -> %dx%d per step (a full 4K frame is %.0fx this)' % (CPU_SAMPLE, CPU_SAMPLE, CPU_SAMPLE * SCALE, CPU_SAMPLE * SCALE,
                                                                               H_IN * W_IN / CPU_SAMPLE ** 2)
  print(json.dumps({
    'impl': 'reference', 'metric': METRIC, 'value': val, 'unit': 'MPix/s', 'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
    'ms_per_step': sec * 1e3, 'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
    'config': {'workload': 'a4 4x SR, bounded sample of the 3840x2160 frame', 'weights': wsrc, 'sample': sample},
    'cpu_baseline': {'value': val, 'unit': 'MPix/s', 'cores': threads, 'kind': 'port', 'sample': sample},
    'e2e': {'value': val, 'unit': 'MPix/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    'gpu_launches': 0}))


# ------------------------------------------------------------------------------------------------
def main():
  ap = argparse.ArgumentParser()
  ap.add_argument('--gpus', type=int, default=1)
  ap.add_argument('--steps', type=int, default=5)
  ap.add_argument('--warmup', type=int, default=3)
  ap.add_argument('--impl', default='native', choices=['native', 'reference'])
  ap.add_argument('--no-cpu-baseline', action='store_true')
  args = ap.parse_args()
  rank = int(os.environ.get('RANK', '0'))
  world = int(os.environ.get('WORLD_SIZE', '1'))
  local = int(os.environ.get('LOCAL_RANK', '0'))
  if args.impl == 'reference':
    return run_reference(args, rank)
  # stdout carries exactly ONE JSON line: native libraries (NCCL prints its version banner on fd 1 when the first
  # communicator is created) are sent to stderr for the duration of the run
  sys.stdout.flush()
  real_stdout = os.dup(1)
  os.dup2(2, 1)

  def emit(line):
    sys.stdout.flush()
    os.dup2(real_stdout, 1)
    print(json.dumps(line), flush=True)

  import torch
  import torch.distributed as dist
  from moephoto_b200 import runSR, imageProcess as IP, parallel as PAR, _lib
  from moephoto_b200.config import config
  import ctypes

  if not torch.cuda.is_available():
    raise SystemExit('bench.py: no GPU — the engine has no CPU fallback (use --impl reference for the CPU leg)')
  args.warmup = max(3, args.warmup)
  torch.cuda.set_device(local)
  config.deviceId = local
  if world > 1:
    # the image exports NCCL_DEBUG=VERSION, which makes NCCL print its banner on STDOUT in front of the JSON line
    if os.environ.get('NCCL_DEBUG', '').upper() in ('', 'VERSION'):
      os.environ['NCCL_DEBUG'] = 'WARN'
    os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
  dev = torch.device('cuda', local)
  eng = IP.getEngine(local)
  sd, wsrc = a4_weights()
  # the reference's auto plan for a 180 GB device: pinned so every rank and every run agree
  config.freeMemOverride = int(178 * 2 ** 30 * .9)
  opt = runSR.getOpt({'model': 'a', 'scale': SCALE}, weights=sd)

  frame_u8 = synthetic_frame(H_IN, W_IN, 0)
  host_in = torch.from_numpy(frame_u8).pin_memory()
  host_out = torch.empty((H_IN * SCALE, W_IN * SCALE, 3), dtype=torch.uint8).pin_memory() if rank == 0 else None
  x = IP.toTorch(8)(frame_u8) if rank == 0 else torch.empty((3, H_IN, W_IN), dtype=torch.half, device=dev)

  def step():
    if world == 1:
      return runSR.sr(opt)(x)
    return PAR.sharded_doCrop(opt, x)

  def sync():
    torch.cuda.synchronize()
    if world > 1:
      dist.barrier()
      torch.cuda.synchronize()

  def timed(fn, k):
    sync()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(k):
      r = fn()
    b.record()
    sync()
    ms = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device=dev)
    if world > 1:
      dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return ms.item(), r

  sampler = ClockSampler(local) if rank == 0 else None   # started before the warm-up: nvidia-smi needs ~1 s to begin sampling
  for _ in range(args.warmup):
    y = step()
  sync()
  plan = opt.plan
  l0 = eng.launches()
  if sampler:
    sampler.mark()
  eng.profile(True)
  total_ms, y = timed(step, args.steps)
  eng.profile(False)
  prof = eng.profile_read()
  clocks = sampler.stop() if sampler else None
  launches = eng.launches() - l0
  ms_step = total_ms / args.steps
  out_mpix = H_IN * SCALE * W_IN * SCALE / 1e6
  value = out_mpix / (ms_step / 1e3)

  # ---- e2e: host uint8 frame -> host uint8 result, copies inside the timed region
  def e2e_step():
    if world == 1:
      _lib.check(eng.lib.moe_enhance_host(opt.modelCached.handle, ctypes.c_void_p(host_in.data_ptr()), 8, ctypes.byref(plan.c),
                                          ctypes.c_void_p(host_out.data_ptr()), 8, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
      return None
    PAR.sharded_enhance_host(opt, host_in.numpy() if rank == 0 else None, shared_out, 8, 8)
    return None

  del y
  shared_out = None
  if world > 1:
    # the result frame lives in a /dev/shm mapping every rank page-locks: each GPU writes its own band over its own PCIe link
    name = 'moephoto_b200_bench_%s' % os.environ.get('MASTER_PORT', '0')
    if rank == 0:
      shared_out = PAR.SharedHostFrame(name, (H_IN * SCALE, W_IN * SCALE, 3), torch.uint8, create=True)
    dist.barrier()
    if rank != 0:
      shared_out = PAR.SharedHostFrame(name, (H_IN * SCALE, W_IN * SCALE, 3), torch.uint8)
  for _ in range(2):
    e2e_step()
  e2e_ms, _ = timed(e2e_step, args.steps)
  e2e_val = out_mpix / (e2e_ms / args.steps / 1e3)

  if shared_out is not None:
    dist.barrier()
    shared_out.close(unlink=(rank == 0))
  if rank != 0:
    if world > 1:
      dist.barrier()
      dist.destroy_process_group()
    return

  pk = peaks()
  conv_ms, conv_flops, conv_n = prof['conv3x3']
  achieved = conv_flops / (conv_ms * 1e-3) / 1e12 if conv_ms > 0 else 0.0
  traffic = None
  tpath = os.path.join(ROOT, 'profiles', 'traffic.json')
  if os.path.exists(tpath):
    traffic = json.load(open(tpath)).get('conv3x3_tc_kernel_dram_bytes_per_launch')
  line = {
    'metric': METRIC, 'value': value, 'unit': 'MPix/s', 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
    'ms_per_step': ms_step, 'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f16 storage / f32 accumulate',
    'data': 'synthetic',
    'config': {'workload': 'a4 (models.Net4x) 4x SR, one 3840x2160 RGB frame -> 15360x8640 (BASELINE configs[2])', 'weights': wsrc,
               'tile_plan': '%d reference tiles %s, pad 5, seam %d px' % (len(plan.tiles), [(t[1] - t[0], t[3] - t[2]) for t in plan.tiles][:4], plan.pad_sc),
               'parallelism': 'rows of the canvas sharded over %d GPU(s), 16-px recompute halo' % world,
               'l2': 'per-layer activations are 0.8-51 GB per tile, far larger than the 126 MB L2; no flush needed'},
    'clocks': clocks,
    'e2e': {'value': e2e_val, 'unit': 'MPix/s', 'ms_per_step': e2e_ms / args.steps, 'h2d_bytes_per_step': H_IN * W_IN * 3,
            'd2h_bytes_per_step': H_IN * SCALE * W_IN * SCALE * 3, 'path': 'moe_enhance_host (C ABI, pinned host uint8 in/out)' if world == 1 else
            'toTorch on rank 0 -> NCCL broadcast -> per-rank row band -> moe_to_output + D2H of each band into a shared page-locked host frame'},
    'gpu_launches': launches,
    'roofline': {'bound': 'tensor', 'kernel': 'conv3x3_pair_trunk_kernel + conv3x3_pair_kernel + conv3x3_pair_head_kernel (all %d 3x3-convolution launches of rank 0 in the timed region)' % conv_n,
                 'achieved': achieved, 'peak': pk['tflops'], 'unit': 'TFLOP/s', 'frac': achieved / pk['tflops'], 'peak_source': pk['src'],
                 'traffic': traffic, 'avg_launch_ms': conv_ms / max(1, conv_n),
                 'algorithmic_flops_per_launch': conv_flops / max(1, conv_n),
                 'share_of_step': conv_ms / total_ms,
                 'other_kernels': {k: {'ms_per_step': v[0] / args.steps, 'GBps': (v[1] / (v[0] * 1e-3) / 1e9 if v[0] > 0 else 0.0), 'launches': v[2]}
                                   for k, v in prof.items() if k in ('conv_input', 'head')},
                 'conv_breakdown': {k: {'ms_per_step': v[0] / args.steps, 'TFLOPs': (v[1] / (v[0] * 1e-3) / 1e12 if v[0] > 0 else 0.0), 'launches': v[2]}
                                    for k, v in prof.items() if k in ('conv_trunk', 'conv_up')}},
    'whole_step_tflops': FLOP_PER_LR_PIXEL_PLANE * 3.0 * H_IN * W_IN / (ms_step * 1e-3) / 1e12,
  }
  if world == 1 and not args.no_cpu_baseline:
    import torch as _t
    threads = host_threads()
    _t.set_num_threads(threads)
    crop = synthetic_frame(CPU_SAMPLE, CPU_SAMPLE, 1)
    cpu_port_step(sd, crop, threads)
    ts = [cpu_port_step(sd, crop, threads) for _ in range(3)]
    sec = min(ts)
    line['cpu_baseline'] = {'value': (CPU_SAMPLE * SCALE) ** 2 / 1e6 / sec, 'unit': 'MPix/s', 'cores': threads, 'kind': 'port',
                            'sample': 'a4 on a %dx%d crop (1 warm-up, best of 3), oracle port with PyTorch CPU conv2d, fp32' % (CPU_SAMPLE, CPU_SAMPLE)}
  emit(line)
  if world > 1:
    dist.barrier()
    dist.destroy_process_group()


if __name__ == '__main__':
  main()
