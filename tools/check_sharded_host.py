"""torchrun --nproc-per-node N tools/check_sharded_host.py : the multi-GPU paths of moephoto_b200/parallel.py must give the
bytes of the single-GPU chain toTorch -> doCrop -> toOutput:
  * BandSharder.run / run_host (peer memory: every rank reads the root's frame and stores into the root's canvas in place),
  * sharded_enhance_host (NCCL broadcast; every rank copies its own band into a shared page-locked host frame)."""
import os, sys
import numpy as np
import torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import helpers as H
from moephoto_b200 import runSR, imageProcess as IP, parallel as PAR
from moephoto_b200.config import config

rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local); config.deviceId = local
os.environ['NCCL_DEBUG'] = 'WARN'
dist.init_process_group('nccl', device_id=torch.device('cuda', local))
config.freeMemOverride, config.crop_sr = int(4e9), 96
opt = runSR.getOpt({'model': 'a', 'scale': 4}, weights=H.load_weights('a4'))
img = np.random.default_rng(3).integers(0, 256, (150, 260, 3), dtype=np.uint8)
name = 'moephoto_b200_check_%s' % os.environ.get('MASTER_PORT', '0')
if rank == 0:
  out = PAR.SharedHostFrame(name, (600, 1040, 3), torch.uint8, create=True)
dist.barrier()
if rank != 0:
  out = PAR.SharedHostFrame(name, (600, 1040, 3), torch.uint8)
for _ in range(2):
  PAR.sharded_enhance_host(opt, img if rank == 0 else None, out, 8, 8)
if rank == 0:
  ref_opt = runSR.getOpt({'model': 'a', 'scale': 4}, weights=H.load_weights('a4'))
  xr = IP.toTorch(8)(img)
  want16 = IP.doCrop(ref_opt, xr)
  want = IP.toOutput(8)(want16)
  got = out.tensor.numpy()
  print('sharded_enhance_host world=%d: identical=%s registered=%s tiles=%d' % (world, bool(np.array_equal(got, want)), out._registered, len(opt.plan.tiles)), flush=True)
  out.tensor.zero_()
dist.barrier()
sh = PAR.BandSharder(opt, (3, 150, 260), torch.device('cuda', local))
for i in range(3):
  y = sh.run(xr if rank == 0 else None)
  torch.cuda.synchronize()
  if rank == 0:
    print('BandSharder.run world=%d pass %d: peer memory=%s identical=%s band rows=%s' % (world, i, sh.peer, bool(torch.equal(y, want16)), (sh.lo, sh.hi)), flush=True)
    y.zero_()
  dist.barrier()
for i in range(2):
  sh.run_host(img if rank == 0 else None, out, 8, 8)
  if rank == 0:
    print('BandSharder.run_host world=%d pass %d: identical=%s' % (world, i, bool(np.array_equal(out.tensor.numpy(), want))), flush=True)
    out.tensor.zero_()
  dist.barrier()
out.close(unlink=(rank == 0))
dist.destroy_process_group()
