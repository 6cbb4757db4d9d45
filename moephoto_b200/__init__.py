"""moephoto_b200 — Blackwell-native (sm_100a) engine for MoePhoto's tiled SR / denoise hot path.

Host-side mirror of the reference's operator interface for this path:
  moephoto_b200.runSR        (getOpt, sr, mode_switch, ramCoef)          <- python/runSR.py
  moephoto_b200.runDN        (getOpt, mode_switch, ramCoef)              <- python/runDN.py
  moephoto_b200.imageProcess (Option, initModel, prepare, doCrop, ...)   <- python/imageProcess.py
  moephoto_b200.parallel     row-band sharding of one image over the GPUs of a box (new; the reference
                             is single-GPU)
All compute goes through the C ABI in include/moephoto_b200.h (moephoto_b200/lib/libmoephoto_b200.so).
"""
__version__ = '0.1.0'
