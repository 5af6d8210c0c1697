"""vipnerf_b200 - B200-native (sm_100a) volumetric render path for ViP-NeRF.

    from vipnerf_b200.ModelFactory import get_model      # reference-compatible plugin factory
    from vipnerf_b200 import renderpath                   # tensor-level wrappers of the C ABI

The CUDA shared library (libvipnerf_b200.so, C ABI in include/vipnerf.h) is built in-tree by
`python -m vipnerf_b200.build`; nothing here falls back to PyTorch or the CPU.
"""
__version__ = '0.1.0'
