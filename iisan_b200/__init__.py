"""iisan_b200 -- B200-native (sm_100a) implementation of the IISAN(Cached) / IISAN-Versa training hot path.

Layout: csrc/ (hand-written CUDA + the C ABI of include/iisan_b200.h), _lib.py (ctypes binding),
ops.py (autograd glue), plan.py (host logic), model/ and model_asym/ (mirrors of the reference's
`model` package), parallel.py (data-parallel helpers), store.py (cached hidden-state store).
"""
__version__ = "0.1.0"
