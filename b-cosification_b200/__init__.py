"""bcos_b200 -- B200-native (sm_100a) B-cos forward + dynamic-linear explanation path.

Layout
  csrc/      hand-written CUDA kernels + the C-ABI (`libbcosk.so`, declared in include/bcosk.h)
  _lib.py    ctypes binding of the C-ABI (fails loudly when the library is missing)
  modules/   drop-in mirror of the reference's `bcos.modules` surface
  engine/    fused execution plans (conv+BN+ReLU+residual epilogues, explain dgrad chain)
  models.py  offline builders for the B-cosified networks (reference state-dict key names)
  utils/     synthetic workload (weights, images), metrics
"""
__version__ = "0.1.0"
