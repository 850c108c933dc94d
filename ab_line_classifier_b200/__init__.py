"""B200-native (sm_100a) numerics for the cutoffvgg16 hot path of gchetty/ab-line-classifier.

Host side is Python (the reference is Python); the numerics are hand-written CUDA kernels in
``csrc/`` behind the C ABI of ``include/abl_b200.h`` (``libabl_b200.so``), bound with ctypes.
"""
__version__ = "0.1.0"
