"""android_svo_b200 — B200-native (sm_100a) implementation of SVO's semi-direct tracking front end
behind the reference's operator surface.  The compute path is libsvob200.so (hand-written CUDA
kernels behind the C ABI in include/svob200.h); this package only holds what that path needs:

  csrc/      CUDA kernels + the C ABI
  host/      C++ host-side mirror of the reference's operator interface (svo::SparseImgAlign::run, ...)
  capi.py    ctypes bindings used by tests/ and bench.py
  synth.py   deterministic synthetic sequences
  frontend.py  per-frame front-end pipeline (pyramid -> align -> refine -> seed update) over the C ABI

There is no CPU fallback anywhere in this package.
"""
__all__ = ["capi", "synth"]
