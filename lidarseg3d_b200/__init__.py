"""lidarseg3d_b200 - B200-native (sm_100a) MSeg3D / SDSeg3D forward path behind the det3d plugin API.

Layout: ``csrc/`` hand-written CUDA + C ABI (include/ls3d.h), ``capi.py`` ctypes binding,
``det3d/`` host-side mirror of the reference registry / builder / module interface.
"""
__version__ = "0.1.0"
