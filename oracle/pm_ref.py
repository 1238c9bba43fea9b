"""ctypes front of oracle/_ref/libpm_ref_*.so: the reference's own receive loop (PacketMaster.c:304-397, extracted
and compiled by oracle/build_pm_ref.py).  TEST INFRASTRUCTURE: this is the pin of oracle/decode.py,
oracle/packetmaster_core.c and the GPU decode."""
import ctypes

import numpy as np

from . import build_pm_ref
from .decode import BUFSIZE_INTS, words_to_wire


def available(R, npix, cap=2500):
    return build_pm_ref.build(R, npix, cap) is not None


def run(streams, npix_per_roach, exptime, cap=2500, want_lists=False, filler_adr=254):
    """Feeds every roach stream (flat u64 words, arrival order) to the reference's loop, bundle by bundle in round-robin
    order over the roaches (as the socket loop would, PacketMaster.c:251-277).  Streams that are not a whole number of
    8192-word bundles are padded with "non-pixel" filler words of channel `filler_adr`, whose
    count is taken off the non-pixel statistic again.  Returns the dict of oracle.decode.packetmaster_bin plus
    'sec' [R], and with want_lists 'list_len' [exptime][R*npix] and 'lists' (u64 [exptime][R*npix][cap])."""
    R = len(streams)
    so = build_pm_ref.build(R, npix_per_roach, cap)
    if so is None:
        raise RuntimeError('libpm_ref for R=%d npix=%d cap=%d is not built and /root/reference is absent' % (R, npix_per_roach, cap))
    lib = ctypes.CDLL(so)
    assert lib.pm_ref_nroaches() == R and lib.pm_ref_npix() == npix_per_roach
    lib.pm_ref_list_len.restype = ctypes.POINTER(ctypes.c_int32)
    lib.pm_ref_lists.restype = ctypes.POINTER(ctypes.c_uint64)
    lib.pm_ref_init(int(exptime), 1 if want_lists else 0)
    assert filler_adr >= npix_per_roach and filler_adr != 255
    wires, n_pad = [], 0
    for st in streams:
        st = np.asarray(st, dtype=np.uint64)
        pad = (-st.size) % BUFSIZE_INTS
        if pad:
            st = np.concatenate([st, np.full(pad, np.uint64(filler_adr) << np.uint64(56), dtype=np.uint64)])
        wires.append(np.frombuffer(words_to_wire(st), dtype=np.uint32).reshape(-1, 2, BUFSIZE_INTS))
    # the filler after the stream is seen only while sec < exptime (PacketMaster.c:327)
    k = 0
    while any(k < w.shape[0] for w in wires):
        for r, w in enumerate(wires):
            if k < w.shape[0]:
                lo = np.ascontiguousarray(w[k, 0]); hi = np.ascontiguousarray(w[k, 1])
                lib.pm_ref_bundle(r, lo.ctypes.data_as(ctypes.c_void_p), hi.ctypes.data_as(ctypes.c_void_p))
        k += 1
    npt = R * npix_per_roach
    counts = np.zeros((exptime, npt), dtype=np.int32)
    lib.pm_ref_counts(counts.ctypes.data_as(ctypes.c_void_p))
    sec = np.zeros(R, dtype=np.int32)
    lib.pm_ref_sec(sec.ctypes.data_as(ctypes.c_void_p))
    st3 = np.zeros(3, dtype=np.int64)
    lib.pm_ref_stats(st3.ctypes.data_as(ctypes.c_void_p))
    for r, st in enumerate(streams):          # filler words seen while the roach's seconds were still open
        st = np.asarray(st, dtype=np.uint64)
        pad = (-st.size) % BUFSIZE_INTS
        if pad:
            n_eos = int(((st >> np.uint64(56)) == 255).sum())
            if n_eos < exptime:
                n_pad += pad
    out = dict(counts=counts.astype(np.int64), sec=sec, n_eos=int(st3[0]), n_corrupt_eos=int(st3[1]),
               n_nonpixel=int(st3[2]) - n_pad)
    if want_lists:
        ln = np.ctypeslib.as_array(lib.pm_ref_list_len(), shape=(exptime, npt)).copy()
        ls = np.ctypeslib.as_array(lib.pm_ref_lists(), shape=(exptime, npt, cap)).copy()
        out['list_len'], out['lists'] = ln, ls
    lib.pm_ref_free()
    return out
