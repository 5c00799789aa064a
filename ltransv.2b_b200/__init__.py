"""B200-native drop-in for the LTRANS v.2b per-particle time step.

Only what the hot path needs lives here:
  csrc/   hand-written sm_100a CUDA kernels + the C ABI (include/ltrans_b200.h)
  host/   host-side mirror of the reference interface (namelist, CSV formats,
          run loop, ctypes binding of the C ABI, synthetic ROMS world)
The CUDA library is mandatory: there is no CPU fallback (see host/binding.py).
"""
__all__ = ["host"]
