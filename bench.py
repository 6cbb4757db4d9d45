"""bench.py — the BASELINE.json metric on the BASELINE.json config.

  metric : output MPix/s of 4x SR (a4 = models.Net4x) on one 3840x2160 RGB frame -> 15360x8640
  step   : one frame through runSR.sr(getOpt({'model':'a','scale':4})) — the reference's auto tile plan
           for a 180 GB GPU (4 column strips of 2160x968, pad 5, seam 20 px), every tile through the 20
           convolutions of Net4x, seam-blended and stitched on the device.
  N GPUs : the SAME frame, canvas rows sharded over the ranks (moephoto_b200/parallel.py):
           per-rank row band + 16-px recompute halo; every rank reads the frame from and stores its band into
           rank 0's HBM in place over NVLink peer memory (parallel.BandSharder; no bulk collective) -> "strong".
  value  : frame resident in HBM (fp16 planar) -> stitched fp16 canvas resident in HBM on rank 0; the engine's per-launch
           profiler is OFF in this timed region (the per-kernel breakdown comes from a second pass).
  e2e    : uint8 HWC frame in pinned HOST memory -> uint8 HWC result in HOST memory, through the C-ABI
           (moe_enhance_host at N=1; at N>1 parallel.BandSharder.run_host: the uint8 frame is uploaded once into rank 0's HBM,
           every rank reads it over peer memory, computes its band and copies it into a shared page-locked host frame over its
           own PCIe link, moe_run_band_to_host), copies timed.
  roofline : the dominant kernel (conv3x3_pair_head_kernel) and, under `kernels`, every kernel class with its own algorithmic
           work (halo rows, padded channels and scrap columns are NOT counted), CUDA-event time, fraction of
           the measured peak and ncu DRAM bytes per launch (profiles/traffic.json, one capture session of this build).
  configs  : the other named configurations of BASELINE.json (C1 256x256 a2, C2 1080p a2, C4 dn_lite15 -> a2 on 16 x 1080p,
           C5 a3 on 64 x 4K, frames sharded over the ranks) measured in the same run, after the headline.
  --impl reference : the reference's CPU path (PyTorch conv2d on the host cores, fp32) restated by the
           oracle port (oracle/net.py forward_torch + oracle/tiling.py), all host threads, on a
           bounded sample of the same workload (a 512x512 crop per step).

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
CONV_FLOP_PER_LR_PIXEL_PLANE = 3907584     # its 3x3 convolutions 64 -> 64 / 256 (conv_input and the two heads excluded)
CPU_SAMPLE = int(os.environ.get('MOE_BENCH_CPU_SAMPLE', '512'))   # the CPU legs run a CPU_SAMPLE^2 crop per step


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
  """SM clock and throttle reasons DURING the timed region (B200_PROFILING.md recipe).  NVML is polled from a thread every 5 ms
  (the 8-GPU timed region lasts ~120 ms: `nvidia-smi -lms 50` delivered no sample inside it); `nvidia-smi` is the fallback when
  the NVML binding is missing."""
  Q = 'index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,' \
      'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

  def __init__(self, gpu_index):
    self.rows, self.skip, self.p, self.f, self.thread = [], 0, None, None, None
    try:
      import pynvml, threading
      pynvml.nvmlInit()
      try:
        uuid = str(torch.cuda.get_device_properties(gpu_index).uuid)
        h = pynvml.nvmlDeviceGetHandleByUUID(('GPU-' + uuid) if not uuid.startswith('GPU-') else uuid)
      except Exception:
        h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
      self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
      bits = {'hw_slowdown': pynvml.nvmlClocksEventReasonHwSlowdown, 'hw_thermal_slowdown': pynvml.nvmlClocksEventReasonHwThermalSlowdown,
              'sw_thermal_slowdown': pynvml.nvmlClocksEventReasonSwThermalSlowdown, 'sw_power_cap': pynvml.nvmlClocksEventReasonSwPowerCap}
      self.running = True

      def poll():
        while self.running:
          try:
            mhz = float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
            mask = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
            self.rows.append((mhz, [k for k, b in bits.items() if mask & b]))
          except Exception:
            pass
          time.sleep(0.005)
      self.thread = threading.Thread(target=poll, daemon=True)
      self.thread.start()
      self.how = 'NVML polled every 5 ms'
      return
    except Exception:
      self.thread = None
    self.how = 'nvidia-smi -lms 50'
    self.f = tempfile.NamedTemporaryFile('w+', suffix='.csv', delete=False)
    try:
      self.p = subprocess.Popen(['nvidia-smi', '-i', str(gpu_index), '--query-gpu=' + self.Q, '--format=csv,noheader,nounits', '-lms', '50'],
                                stdout=self.f, stderr=subprocess.DEVNULL)
    except Exception:
      pass

  def mark(self):
    """samples taken before this call belong to the warm-up and are dropped"""
    if self.thread:
      self.skip = len(self.rows)
      return
    self.f.flush()
    try:
      self.skip = sum(1 for _ in open(self.f.name))
    except Exception:
      self.skip = 0

  def stop(self):
    out = dict(sm_mhz=None, sm_max_mhz=None, reasons=[], samples=0, how=self.how)
    sm, reasons = [], set()
    if self.thread:
      self.running = False
      self.thread.join(timeout=2)
      out['sm_max_mhz'] = self.max_mhz
      for mhz, rs in self.rows[self.skip:]:
        sm.append(mhz); reasons.update(rs)
    elif self.p is not None:
      self.p.terminate()
      try:
        self.p.wait(timeout=5)
      except Exception:
        self.p.kill()
      self.f.flush()
      rows = [r.strip().split(', ') for r in open(self.f.name) if r.strip()][self.skip:]
      os.unlink(self.f.name)
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
# what each kernel class of the engine profiler is, for the roofline report: (kernel, bound, unit of the achieved figure)
KERNELS = {
  'conv_up_head': ('conv3x3_pair_head_kernel (last upsample conv 64->256 + PixelShuffle + PReLU fused with the heads\' dot products)', 'tensor'),
  'conv_up': ('conv3x3_pair_kernel (first upsample conv 64->256 + PixelShuffle + PReLU)', 'tensor'),
  'arsb': ('arsb_pair_kernel (one residual block: conv_1 + PReLU + conv_2 + x scale + skip)', 'tensor'),
  'conv_trunk': ('conv3x3_pair_trunk_kernel (conv_input2, 64->64)', 'tensor'),
  'conv_input': ('conv_first_kernel (1->64 + PReLU)', 'hbm'),
  'head': ('head_stencil_kernel (vertical third of both head stencils, branch sum, seam blend, canvas store)', 'hbm'),
}


def kernel_rooflines(prof, steps, step_ms, pk, traffic, work_scale=1.0):
  """per kernel class: time share, achieved algorithmic rate, fraction of the measured peak, ncu DRAM bytes per launch"""
  out = {}
  for cls, (name, bound) in KERNELS.items():
    ms, work, n = prof[cls]
    if n == 0 or ms <= 0:
      continue
    rate = work * work_scale / (ms * 1e-3)
    peak = pk['tflops'] if bound == 'tensor' else pk['hbm']
    achieved = rate / 1e12 if bound == 'tensor' else rate / 1e9
    t = (traffic or {}).get(cls)
    out[cls] = {'kernel': name, 'bound': bound, 'ms_per_step': ms / steps, 'share_of_step': ms / steps / step_ms, 'launches_per_step': n / steps,
                'achieved': achieved, 'peak': peak, 'unit': 'TFLOP/s' if bound == 'tensor' else 'GB/s', 'frac': achieved / peak,
                'algorithmic_work_per_launch': work * work_scale / n, 'traffic': t}
  return out


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument('--gpus', type=int, default=1)
  ap.add_argument('--steps', type=int, default=5)
  ap.add_argument('--warmup', type=int, default=3)
  ap.add_argument('--impl', default='native', choices=['native', 'reference'])
  ap.add_argument('--no-cpu-baseline', action='store_true')
  ap.add_argument('--no-extras', action='store_true', help='skip the other BASELINE configs, the parity figures and the eager-GPU baseline')
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
  sharder = PAR.BandSharder(opt, (3, H_IN, W_IN), dev) if world > 1 else None

  def step():
    if world == 1:
      return runSR.sr(opt)(x)
    return sharder.run(x)

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
  total_ms, y = timed(step, args.steps)                   # THE timed region: profiler off
  clocks = sampler.stop() if sampler else None
  launches = eng.launches() - l0
  ms_step = total_ms / args.steps
  out_mpix = H_IN * SCALE * W_IN * SCALE / 1e6
  value = out_mpix / (ms_step / 1e3)

  # ---- second pass, same steps, per-launch CUDA events on: the per-kernel breakdown (not part of `value`)
  prof_steps = min(args.steps, 3)
  eng.profile(True)
  prof_total_ms, y = timed(step, prof_steps)
  eng.profile(False)
  prof = eng.profile_read()

  # ---- e2e: host uint8 frame -> host uint8 result, copies inside the timed region
  def e2e_step():
    if world == 1:
      _lib.check(eng.lib.moe_enhance_host(opt.modelCached.handle, ctypes.c_void_p(host_in.data_ptr()), 8, ctypes.byref(plan.c),
                                          ctypes.c_void_p(host_out.data_ptr()), 8, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
      return None
    sharder.run_host(host_in.numpy() if rank == 0 else None, shared_out, 8, 8)
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

  # ---- the other named configurations (every rank takes part: frames are sharded over the ranks)
  extras = None
  if not args.no_extras:
    extras = other_configs(rank, world, dev, sync)

  if rank != 0:
    if world > 1:
      dist.barrier()
      dist.destroy_process_group()
    return

  pk = peaks()
  traffic = None
  tpath = os.path.join(ROOT, 'profiles', 'traffic.json')
  if os.path.exists(tpath):
    traffic = json.load(open(tpath)).get('dram_bytes_per_launch')
  # algorithmic work never counts halo rows: rank 0's band is in_h / world LR rows (its kernels computed up to 32 more)
  lo, hi = PAR.band_rows(H_IN, 1, world, 0)
  conv_ms, conv_flops, conv_n = prof['conv3x3']
  alg_conv_flops = CONV_FLOP_PER_LR_PIXEL_PLANE * 3.0 * (hi - lo) * W_IN * prof_steps
  work_scale = alg_conv_flops / conv_flops if conv_flops > 0 else 1.0          # 1.0 at N = 1 (no halo), < 1 on a band
  prof_step_ms = prof_total_ms / prof_steps
  kernels = kernel_rooflines(prof, prof_steps, prof_step_ms, pk, traffic if world == 1 else None, work_scale)
  dom = kernels.get('conv_up_head', {})
  achieved_all = alg_conv_flops / (conv_ms * 1e-3) / 1e12 if conv_ms > 0 else 0.0
  line = {
    'metric': METRIC, 'value': value, 'unit': 'MPix/s', 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
    'ms_per_step': ms_step, 'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f16 storage / f32 accumulate',
    'data': 'synthetic',
    'config': {'workload': 'a4 (models.Net4x) 4x SR, one 3840x2160 RGB frame -> 15360x8640 (BASELINE configs[2])', 'weights': wsrc,
               'tile_plan': '%d reference tiles %s, pad 5, seam %d px' % (len(plan.tiles), [(t[1] - t[0], t[3] - t[2]) for t in plan.tiles][:4], plan.pad_sc),
               'parallelism': 'rows of the canvas sharded over %d GPU(s), 16-px recompute halo' % world,
               'l2': 'per-layer activations are 0.8-51 GB per tile, far larger than the 126 MB L2; no flush needed',
               'numerics': 'the rounding points of the reference\'s GPU fp16 path (every aten op rounds; DESIGN.md §3)'},
    'clocks': clocks,
    'e2e': {'value': e2e_val, 'unit': 'MPix/s', 'ms_per_step': e2e_ms / args.steps, 'h2d_bytes_per_step': H_IN * W_IN * 3,
            'd2h_bytes_per_step': H_IN * SCALE * W_IN * SCALE * 3, 'path': 'moe_enhance_host (C ABI, pinned host uint8 in/out)' if world == 1 else
            'BandSharder.run_host: H2D of the uint8 frame on rank 0 -> every rank reads it over NVLink peer memory -> per-rank row band -> moe_run_band_to_host: D2H of each band into a shared page-locked host frame over the rank\'s own PCIe link'},
    'gpu_launches': launches,
    'roofline': {'bound': 'tensor', 'kernel': dom.get('kernel'), 'achieved': dom.get('achieved'), 'peak': pk['tflops'], 'unit': 'TFLOP/s',
                 'frac': dom.get('frac'), 'traffic': dom.get('traffic'), 'peak_source': pk['src'],
                 'avg_launch_ms': dom.get('ms_per_step', 0) / max(1, dom.get('launches_per_step', 1)),
                 'algorithmic_flops_per_launch': dom.get('algorithmic_work_per_launch'), 'share_of_step': dom.get('share_of_step'),
                 'measured_in': 'a second pass of %d steps with the engine profiler on (%.2f ms per step; the timed region runs with it off)' % (prof_steps, prof_step_ms),
                 'all_conv3x3': {'achieved': achieved_all, 'frac': achieved_all / pk['tflops'], 'ms_per_step': conv_ms / prof_steps,
                                 'share_of_step': conv_ms / prof_total_ms, 'launches_per_step': conv_n / prof_steps,
                                 'note': 'algorithmic FLOPs of the kept rows only (no halo recompute, no padded channels)'},
                 'kernels': kernels},
    'whole_step_tflops': FLOP_PER_LR_PIXEL_PLANE * 3.0 * H_IN * W_IN / (ms_step * 1e-3) / 1e12,
  }
  if extras:
    line.update(extras)
  if world == 1 and not args.no_cpu_baseline:
    import torch as _t
    threads = host_threads()
    _t.set_num_threads(threads)
    crop = synthetic_frame(CPU_SAMPLE, CPU_SAMPLE, 1)
    cpu_port_step(sd, crop, threads)
    ts = [cpu_port_step(sd, crop, threads) for _ in range(3)]
    sec = min(ts)
    line['cpu_baseline'] = {'value': (CPU_SAMPLE * SCALE) ** 2 / 1e6 / sec, 'unit': 'MPix/s', 'cores': threads, 'kind': 'port',
                            'sample': 'a4 on a %dx%d crop (1 warm-up, best of 3; a full 4K frame is %.0fx this, extrapolated by pixel count), oracle port '
                                      'with PyTorch CPU conv2d, fp32' % (CPU_SAMPLE, CPU_SAMPLE, H_IN * W_IN / CPU_SAMPLE ** 2)}
  emit(line)
  if world > 1:
    dist.barrier()
    dist.destroy_process_group()


def other_configs(rank, world, dev, sync):
  """BASELINE.json's other configurations, measured after the headline on the same GPUs (device-timed where the data is resident,
  wall-clock over host-to-host frame streams; max over ranks).  Every rank calls this; rank 0 returns the report.
    C1 256x256 a2 single tile            C2 1920x1080 -> 3840x2160 a2
    C4 dn_lite15 -> a2 on 16 x 1080p     C5 a3 on 64 x 4K frames (frames dealt round-robin to the ranks, video.FrameBatcher)
  plus, on rank 0 at N = 1: engine-vs-reference parity on a golden and the reference's GPU arithmetic (PyTorch eager, cuDNN
  half) timed on the same device."""
  import torch
  import torch.distributed as dist
  import helpers as H
  from moephoto_b200 import runSR, runDN, imageProcess as IP, video
  from moephoto_b200.config import config

  def dev_timed(fn, reps):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
      fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps

  def allmax(v):
    t = torch.tensor([v], dtype=torch.float64, device=dev)
    if world > 1:
      dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()

  rep = {}
  o2 = runSR.getOpt({'model': 'a', 'scale': 2}, weights=H.load_weights('a2'))
  if rank == 0:
    x = IP.toTorch(8)(synthetic_frame(256, 256, 1))
    ms = dev_timed(lambda: runSR.sr(o2)(x), 10)
    rep['C1'] = {'what': 'a2 on one 256x256 RGB tile -> 512x512, 1 GPU, device-resident', 'ms': ms, 'MPix/s': 0.262144 / ms * 1e3}
    x = IP.toTorch(8)(synthetic_frame(1080, 1920, 2))
    ms = dev_timed(lambda: runSR.sr(o2)(x), 5)
    rep['C2'] = {'what': 'a2 1920x1080 -> 3840x2160 (%d reference tile), 1 GPU, device-resident' % len(o2.plan.tiles), 'ms': ms,
                 'MPix/s': 8.2944 / ms * 1e3, 'frac_of_conv_flop_roofline': None}
  # C4 / C5: host frames in, host frames out, frames sharded over the ranks
  odn = runDN.getOpt({'model': 'lite15'}, weights=H.load_weights('dn_lite15'))
  o3 = runSR.getOpt({'model': 'a', 'scale': 3}, weights=H.load_weights('a3'))
  for key, opts, (h, w), n_frames, batch, out_mpix, what in (
      ('C4', [odn, o2], (1080, 1920), 16, 8, 8.2944, 'dn_lite15 -> a2 chained on 16 x 1080p frames'),
      ('C5', [o3], (2160, 3840), 64, 2, 74.6496, 'a3 3x SR on 64 x 3840x2160 frames (video path)')):
    base = [synthetic_frame(h, w, 10 + i) for i in range(2)]
    frames = [base[i % 2] for i in range(n_frames)]                        # frame CONTENT repeats, every frame is processed
    mine = len([i for i in range(n_frames) if i % world == rank])
    fb = video.FrameBatcher(opts, h, w, bit_depth=8, swap_rb=False, batch=min(batch, max(1, mine)))
    for _ in fb.process(frames[:2 * world * fb.batch], rank, world, copy=False):   # warm-up: plans, workspaces, pinned staging
      pass
    sync()
    t0 = time.perf_counter()
    done = sum(1 for _ in fb.process(frames, rank, world, copy=False))
    torch.cuda.synchronize()
    sec = allmax(time.perf_counter() - t0)
    assert done == mine
    if rank == 0:
      rep[key] = {'what': what + ', host uint8 frames in / out, frames dealt round-robin to %d GPU(s), %d frame(s) per engine call' % (world, fb.batch),
                  'seconds': sec, 'frames_per_s': n_frames / sec, 'frames_per_s_per_gpu': n_frames / sec / world, 'MPix/s': n_frames * out_mpix / sec}
    del fb
    sync()
  if rank != 0:
    return None
  out = {'configs': rep}
  if world == 1:
    # parity in the same run: the a4 golden through the public API against the reference's own fp16 output and the oracle
    c = H.load_case('a4_single')
    yg = H.run_case_engine(c)
    yc = H.run_case_engine(c, cpu_bias=True)
    orc = H.run_case_oracle(c, mode='ref16')
    d1, d2 = np.abs(yc - c['ref16']), np.abs(yg - orc)
    out['parity'] = {'case': 'a4_single golden (tests/golden/cases.npz, the UNMODIFIED reference in its GPU fp16 configuration, executed on CPU)',
                     'engine_cpu_bias_mode_vs_reference_fp16': {'max_abs': float(d1.max()), 'psnr_db': H.psnr(yc, c['ref16'])},
                     'engine_default_vs_oracle_ref16': {'max_abs': float(d2.max()), 'psnr_db': H.psnr(yg, orc)},
                     'engine_default_vs_reference_fp32_psnr_db': H.psnr(yg, c['ref'])}
    # the reference's GPU arithmetic on this device: PyTorch eager, cuDNN half, op for op what the reference's nn.Modules
    # launch (oracle port; /root/reference does not exist on the GPU box), one reference tile of the 4K frame
    try:
      from oracle import net as ONet
      sd, _ = a4_weights()
      xt = torch.rand(3, 1, H_IN, 968, device=dev).half()
      ONet.forward_torch(sd, xt[:, :, :64], dtype='float16', device='cuda')
      torch.cuda.synchronize()
      t0 = time.perf_counter()
      ONet.forward_torch(sd, xt, dtype='float16', device='cuda')
      torch.cuda.synchronize()
      sec = time.perf_counter() - t0
      out['gpu_eager_baseline'] = {'what': 'PyTorch eager fp16 (cuDNN) port of Net4x.forward on one 3 x 2160 x 968 reference tile incl. the copy of the '
                                           'result to the host that the port ends with; a frame is 4 such tiles + stitching', 'seconds_per_tile': sec,
                                   'MPix/s': 3 * 0 + (H_IN * SCALE * 968 * SCALE) / 1e6 / sec, 'kind': 'port'}
    except Exception as ex:                                             # never let a baseline leg take the bench line down
      out['gpu_eager_baseline'] = {'unavailable': repr(ex)[:200]}
    finally:
      torch.cuda.empty_cache()
  return out


if __name__ == '__main__':
  main()
