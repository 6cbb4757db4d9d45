"""Loader / builder of the C-ABI shared library (include/moephoto_b200.h).

The library is compiled IN-TREE by build() (nvcc, sm_100a only) into moephoto_b200/lib/ and bound with
ctypes — no torch extension: torch only supplies device pointers and the current CUDA stream.
There is no CPU fallback: load() raises if the library is missing, and every compute entry point
returns MOE_ERR_NO_DEVICE (-> RuntimeError) without an sm_100 GPU.
"""
import ctypes
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
# MOE_B200_LIB: load another build of the same ABI (A/B timing of two builds inside one GPU session, tools/)
LIB_PATH = os.environ.get('MOE_B200_LIB') or os.path.join(HERE, 'lib', 'libmoephoto_b200.so')
SOURCES = [os.path.join(HERE, 'csrc', 'engine.cu')]
HEADERS = [os.path.join(HERE, 'csrc', n) for n in sorted(os.listdir(os.path.join(HERE, 'csrc'))) if n.endswith(('.cuh', '.h'))] + \
          [os.path.join(HERE, '..', 'include', 'moephoto_b200.h')]
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
              '-shared', '-Xcompiler', '-fPIC']

MOE_OK, MOE_ERR_INVALID, MOE_ERR_NO_DEVICE, MOE_ERR_NOMEM, MOE_ERR_CUDA = 0, -1, -2, -3, -4


class MoeTile(ctypes.Structure):
  _fields_ = [(n, ctypes.c_int32) for n in ('top', 'bottom', 'left', 'right', 'top_t', 'left_t', 'bsc', 'rsc')]


class MoePlan(ctypes.Structure):
  _fields_ = [(n, ctypes.c_int32) for n in ('n_tiles', 'scale', 'pad_sc', 'in_h', 'in_w', 'pad_h', 'pad_w', 'out_h', 'out_w')] + \
             [('tiles', ctypes.POINTER(MoeTile)), ('ramp', ctypes.POINTER(ctypes.c_float))]


# every symbol include/moephoto_b200.h declares: name -> (restype, argtypes)
_vp, _i, _i64, _sz, _f = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_size_t, ctypes.c_float
_pp = ctypes.POINTER(ctypes.c_void_p)
_plan = ctypes.POINTER(MoePlan)
SYMBOLS = {
  'moe_abi_version': (_i, []),
  'moe_last_error': (ctypes.c_char_p, []),
  'moe_engine_create': (_i, [_i, _pp]),
  'moe_engine_destroy': (None, [_vp]),
  'moe_engine_launch_count': (_i64, [_vp]),
  'moe_engine_profile': (_i, [_vp, _i]),
  'moe_engine_profile_read': (_i, [_vp, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int64)]),
  'moe_engine_set_conv_path': (_i, [_vp, _i]),
  'moe_engine_debug_buffer': (_i, [_vp, _vp, _sz]),
  'moe_engine_check': (_i, [_vp, _vp]),
  'moe_engine_debug_timeout': (_i, [_vp, ctypes.c_uint64]),
  'moe_model_load': (_i, [_vp, _i, _vp, _sz, _pp]),
  'moe_model_free': (None, [_vp]),
  'moe_model_scale': (_i, [_vp]),
  'moe_plan_workspace_bytes': (_sz, [_vp, _i, _plan, _i, _i]),
  'moe_run_plan': (_i, [_vp, _vp, _i64, _i64, _i, _vp, _i64, _i64, _plan, _i, _i, _vp, _sz, _vp]),
  'moe_conv3x3_c64': (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _f, _vp]),
  'moe_axpby_f16': (_i, [_vp, _vp, _vp, _f, _sz, _vp]),
  'moe_to_planar_f16': (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp, _vp]),
  'moe_to_output': (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp, _vp]),
  'moe_enhance_host': (_i, [_vp, _vp, _i, _plan, _vp, _i, _vp]),
  'moe_enhance_host_c': (_i, [_vp, _vp, _i, _i, _plan, _vp, _i, _vp]),
  'moe_run_band_to_host': (_i, [_vp, _vp, _i64, _i64, _i, _plan, _i, _i, _vp, _i, _vp]),
}

_lib = None


def needs_build():
  if not os.path.exists(LIB_PATH):
    return True
  t = os.path.getmtime(LIB_PATH)
  return any(os.path.exists(s) and os.path.getmtime(s) > t for s in SOURCES + HEADERS)


def build(force=False, verbose=False):
  """nvcc -gencode arch=compute_100a,code=sm_100a ... -> moephoto_b200/lib/libmoephoto_b200.so"""
  if not force and not needs_build():
    return LIB_PATH
  os.makedirs(os.path.dirname(LIB_PATH), exist_ok=True)
  cmd = ['nvcc'] + NVCC_FLAGS + (['-Xptxas', '-v'] if verbose else []) + ['-o', LIB_PATH] + SOURCES
  subprocess.run(cmd, check=True)
  return LIB_PATH


def load():
  global _lib
  if _lib is not None:
    return _lib
  if not os.path.exists(LIB_PATH):
    raise RuntimeError('moephoto_b200: %s is missing — run `python -c "import __graft_entry__ as g; g.build()"` '
                       '(there is no CPU fallback)' % LIB_PATH)
  lib = ctypes.CDLL(LIB_PATH)
  for name, (res, args) in SYMBOLS.items():
    try:
      fn = getattr(lib, name)   # AttributeError here = the .so does not export what the header declares
    except AttributeError:
      if os.environ.get('MOE_B200_LIB'):
        continue                # an older build loaded for A/B timing (tools/ab_bench.sh)
      raise
    fn.restype, fn.argtypes = res, args
  if lib.moe_abi_version() != 1:
    raise RuntimeError('moephoto_b200: ABI version mismatch')
  _lib = lib
  return lib


def check(rc):
  """status code -> the exception class the reference's callers expect (SURVEY.md §8b Errors)."""
  if rc == MOE_OK:
    return
  msg = load().moe_last_error().decode('utf-8', 'replace')
  if rc == MOE_ERR_NOMEM:
    raise MemoryError(msg)
  if rc == MOE_ERR_INVALID:
    raise ValueError(msg)
  raise RuntimeError(msg)
