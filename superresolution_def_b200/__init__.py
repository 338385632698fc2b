"""superresolution_def_b200 — B200-native (sm_100a) hot path of GDev96/SuperResolution_Def.

Only what the hot path needs lives here: `csrc/` (CUDA kernels + the C ABI in include/srk.h),
`_capi.py` (ctypes binding) and the host-side mirrors of the reference's module interface.
"""
__version__ = "0.1.0"
