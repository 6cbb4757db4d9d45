"""TEST INFRASTRUCTURE — not product code.

Imports the UNMODIFIED reference (/root/reference/python) on CPU (fp32, or half through
`gpu_fp16_config`) through five harness-side shims so that golden vectors can be generated and the
oracle restatement can be pinned against the real thing.  /root/reference exists only
in the build container: everything here degrades to `available() == False` elsewhere
(the GPU box), and nothing in `-m gpu` tests, smoke() or bench.py depends on it.

Shims (zero edits to the reference):
  1. stub package `gevent`   (progress.py:4, worker.py:3)
  2. stub package `ailut`    (AiLUT.py:14 via procedure.py:15 -> dehaze.py:9)
  3. TORCH_FORCE_NO_WEIGHTS_ONLY_LOAD=1 (imageProcess.py:306 torch.load of protocol-4 pickles)
  4. writable CWD with `model -> /root/reference/model` (all checkpoint paths are ./model/...)
  5. config.cuda is False without a GPU -> CPU fp32 run type 0.
"""
import os
import sys
import tempfile

REF_ROOT = os.environ.get('MOEPHOTO_REFERENCE', '/root/reference')
_state = {}


def available():
  return os.path.isdir(os.path.join(REF_ROOT, 'python')) and os.path.isdir(os.path.join(REF_ROOT, 'model'))


def _write_stubs(d):
  os.makedirs(os.path.join(d, 'gevent'), exist_ok=True)
  with open(os.path.join(d, 'gevent', '__init__.py'), 'w') as f:
    f.write('def idle(*a, **k): pass\ndef sleep(*a, **k): pass\n'
            'def spawn(f, *a, **k): return f(*a, **k)\ndef spawn_later(t, f, *a, **k): return None\n')
  with open(os.path.join(d, 'gevent', 'event.py'), 'w') as f:
    f.write('class Event:\n  def __init__(self): self._f = False\n  def set(self): self._f = True\n'
            '  def clear(self): self._f = False\n  def is_set(self): return self._f\n  def wait(self, *a): return self._f\n')
  os.makedirs(os.path.join(d, 'ailut'), exist_ok=True)
  with open(os.path.join(d, 'ailut', '__init__.py'), 'w') as f:
    f.write('def ailut_transform(*a, **k): raise NotImplementedError("ailut stub")\n')


def load():
  """Returns a namespace dict with the reference modules runSR, runDN, imageProcess, config, models."""
  if _state:
    return _state
  if not available():
    raise RuntimeError('reference tree not present at ' + REF_ROOT)
  os.environ.setdefault('TORCH_FORCE_NO_WEIGHTS_ONLY_LOAD', '1')
  work = tempfile.mkdtemp(prefix='moeref_')
  stubs = os.path.join(work, 'stubs')
  _write_stubs(stubs)
  os.makedirs(os.path.join(work, '.user'), exist_ok=True)
  os.symlink(os.path.join(REF_ROOT, 'model'), os.path.join(work, 'model'))
  old_cwd = os.getcwd()
  os.chdir(work)  # reference paths are CWD-relative; stay here for model loads
  saved_path = list(sys.path)
  # our own package mirrors the names runSR / runDN / imageProcess inside `moephoto_b200`;
  # the reference's are top-level modules, so there is no clash.
  sys.path[:0] = [stubs, os.path.join(REF_ROOT, 'python')]
  try:
    import torch  # noqa: F401
    import config as ref_config
    import imageProcess as ref_ip
    import models as ref_models
    import runSR as ref_runSR
    import runDN as ref_runDN
  finally:
    sys.path[:] = saved_path + [p for p in (stubs, os.path.join(REF_ROOT, 'python'))]
  _state.update(dict(config=ref_config.config, imageProcess=ref_ip, models=ref_models,
                     runSR=ref_runSR, runDN=ref_runDN, work=work, old_cwd=old_cwd))
  return _state


class gpu_fp16_config:
  """Context manager: make the unmodified reference behave as on its GPU fp16 path while executing on CPU —
  config.dtype() -> torch.half (so castModel does model.half(), imageProcess.py:309-318, and the blend ramp is
  evaluated in half, :109) and config.getRunType() -> 2 (the GPU fp16 row of ramCoef, runSR.py:9 / runDN.py:9, so the
  tile plan is the one a GPU run makes).  Every aten op then rounds to fp16 exactly as on the GPU; only the fp32
  summation order inside conv2d differs from cuDNN's.  Harness-side instance attributes; zero edits to the reference."""

  def __init__(self, enable=True):
    self.enable = enable

  def __enter__(self):
    if self.enable:
      import torch
      cfg = load()['config']
      cfg.dtype = lambda: torch.half
      cfg.getRunType = lambda: 2
    return self

  def __exit__(self, *exc):
    if self.enable:
      cfg = load()['config']
      for name in ('dtype', 'getRunType'):
        if name in vars(cfg):
          delattr(cfg, name)
      # castModel casts the CACHED module in place (imageProcess.py:311-318): after a half run its weights stay
      # rounded to fp16 even when cast back to float.  Drop the cache so a later fp32 call reloads the checkpoint.
      load()['imageProcess'].modelCache.clear()
    return False


def run_sr(x, scale, crop=0, model='a', ensemble=0, half=False):
  """reference runSR.sr(getOpt(...))(x) on CPU, fp32 or (half=True) as its GPU fp16 path; returns (y, plan list, opt)."""
  import torch
  ref = load()
  ref['config'].crop_sr = crop if crop else 'auto'
  with gpu_fp16_config(half):
    opt = ref['runSR'].getOpt({'model': model, 'scale': scale, 'ensemble': ensemble})
    with torch.no_grad():
      y = ref['runSR'].sr(opt)(x.half() if half else x)
  return y, list(opt.iterClip()), opt


def run_dn(x, model='lite15', crop=0, strength=1.0, half=False):
  import torch
  ref = load()
  ref['config'].crop_dn = crop if crop else 'auto'
  with gpu_fp16_config(half):
    opt = ref['runDN'].getOpt({'model': model, 'strength': strength})
    with torch.no_grad():
      y = ref['imageProcess'].RGBFilter(opt)(x.half() if half else x)
  return y, list(opt.iterClip()), opt


def bare_net(key):
  """the cached nn.Module, e.g. key='SRa4' / 'DNlite15' (after a getOpt for it)."""
  return load()['imageProcess'].modelCache[key]


def state_dict(path_rel):
  """fp32 state dict of a reference checkpoint, e.g. 'a4/model_new.pth'."""
  import torch
  os.environ.setdefault('TORCH_FORCE_NO_WEIGHTS_ONLY_LOAD', '1')
  return torch.load(os.path.join(REF_ROOT, 'model', path_rel), map_location='cpu', weights_only=False)
