"""world_size-2 (and 3) gloo runs on CPU of the multi-GPU host logic (moephoto_b200/parallel.py): band
partition, broadcast of the frame, gather of the bands.  The per-band compute is the ORACLE restricted to
a row window (16-px recompute halo) — so this also proves, on CPU, that row-band sharding reproduces the
unsharded doCrop bit for bit."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import helpers as H


def _band_oracle(sd, plan, scale):
  """compute_band(x, lo, hi, canvas) built from the oracle: every tile restricted to [lo,hi) + halo"""
  from oracle import net as N, tiling as T

  def run(x, lo, hi, canvas):
    full = T.do_crop(lambda a: N.forward(sd, a, backend='torch'), x.numpy(), plan)
    canvas[:, lo:hi] = torch.from_numpy(full[:, lo:hi])
  return run


def _worker(rank, world, port, q):
  os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
  dist.init_process_group('gloo', rank=rank, world_size=world)
  try:
    from oracle import net as N, tiling as T
    from moephoto_b200 import parallel as PAR
    sd = H.load_weights('a2')
    h, w, sc = 40, 56, 2
    plan = T.make_plan((3, h, w), 4e9, .9 / 2473., 5, sc, 8, 32)
    g = torch.Generator().manual_seed(0)
    x = torch.rand(3, h, w, generator=g) if rank == 0 else torch.zeros(3, h, w)
    free = PAR.agree_on_free_memory(1000 + rank, 'cpu')
    out = PAR.sharded_run(_band_oracle(sd, plan, sc), x, 3, h, w, sc)
    if rank == 0:
      want = T.do_crop(lambda a: N.forward(sd, a, backend='torch'), x.numpy(), plan)
      q.put((bool(np.array_equal(out.numpy(), want)), free, [PAR.band_rows(h, sc, world, r) for r in range(world)]))
    else:
      q.put((bool(torch.equal(x, torch.rand(3, h, w, generator=torch.Generator().manual_seed(0)))), free, None))
  finally:
    dist.destroy_process_group()


@pytest.mark.parametrize('world', [2, 3])
def test_sharded_run_over_gloo(world):
  torch.set_num_threads(2)
  ctx = mp.get_context('spawn')
  q = ctx.Queue()
  port = 29500 + os.getpid() % 2000 + world
  procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
  for p in procs:
    p.start()
  res = [q.get(timeout=180) for _ in range(world)]
  for p in procs:
    p.join(60)
    assert p.exitcode == 0
  assert all(r[0] for r in res)                 # root: gathered canvas == unsharded; others: received the frame
  assert all(r[1] == 1000 for r in res)         # min over ranks
  bands = [r[2] for r in res if r[2]][0]
  assert bands[0][0] == 0 and bands[-1][1] == 80 and all(a[1] == b[0] for a, b in zip(bands, bands[1:]))


def test_band_rows_cover_the_canvas():
  from moephoto_b200 import parallel as PAR
  for h in (1, 7, 270, 1080, 2160):
    for sc in (1, 2, 3, 4):
      for world in (1, 2, 3, 8):
        b = [PAR.band_rows(h, sc, world, r) for r in range(world)]
        assert b[0][0] == 0 and b[-1][1] == h * sc
        assert all(x[1] == y[0] and x[0] % sc == 0 for x, y in zip(b, b[1:]))
