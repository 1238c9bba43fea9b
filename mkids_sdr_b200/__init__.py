"""mkids_sdr_b200 -- B200-native (sm_100a) hot path of the ARCONS/MKID SDR readout.

Drop-in for the reference's (creanero/MKIDS_SDR) LUT synthesis, software channelizer /
pulse detection and photon-word decode / binning paths.  Python host code mirrors the
reference's call surface and calls hand-written CUDA kernels through the ctypes C ABI in
include/mkidgpu.h.  There is no CPU fallback.
"""
from ._lib import Context, MkidError, default_context  # noqa: F401

__all__ = ['Context', 'MkidError', 'default_context']
