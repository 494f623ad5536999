"""mono_vifi_b200 -- B200-native (sm_100a) implementation of Mono-ViFI's self-supervised training inner loop.

Drop-in for the reference's `layers.py` / `networks/*` module API; the arithmetic runs in hand-written CUDA
behind the C ABI declared in include/monovifi_b200.h (libmonovifi_b200.so, loaded with ctypes).  There is no
CPU fallback: importing the ops without the built library, or calling them on CPU tensors, raises.
"""
__version__ = "0.1.0"
