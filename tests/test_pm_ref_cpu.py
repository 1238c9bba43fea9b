"""Row a16 pinned by the reference ITSELF: the receive loop of PacketMaster.c (:304-397), extracted as text from the
reference tree and compiled by oracle/build_pm_ref.py into oracle/_ref/, against the three restatements the other
tests use (oracle/decode.py literal + vectorised, oracle/packetmaster_core.c)."""
import ctypes
import os

import numpy as np
import pytest

from oracle import decode as odec
from oracle import pm_ref

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ragged_streams(seed, R, npix, secs, per_sec, hot=True, corrupt=True):
    rng = np.random.default_rng(seed)
    streams = []
    for r in range(R):
        parts = []
        for s in range(secs + 1):
            n = int(per_sec * (0.5 + rng.random()))
            ch = rng.integers(0, npix + 2, n)
            if hot:
                ch[rng.random(n) < 0.2] = 1
            w = odec.pack_word(ch, rng.integers(0, 4096, n), rng.integers(0, 4096, n), rng.integers(0, 4096, n),
                               np.sort(rng.integers(0, 10 ** 6, n)))
            eos = np.array([0xFFFFFFFFFFFFFFFF if (s != 1 or r != 0 or not corrupt) else 0xFF00000000000001], dtype=np.uint64)
            parts += [w, eos]
        streams.append(np.concatenate(parts))
    return streams


def _compare(streams, npix, secs, cap):
    if not pm_ref.available(len(streams), npix, cap):
        pytest.skip('oracle/_ref not built (no /root/reference here)')
    ref = pm_ref.run(streams, npix, secs, cap, want_lists=True)
    vec = odec.packetmaster_bin(streams, npix, secs, cap, want_lists=True)
    assert np.array_equal(ref['counts'], vec['counts'])
    for k in ('n_eos', 'n_corrupt_eos', 'n_nonpixel'):
        assert ref[k] == vec[k], (k, ref[k], vec[k])
    # the rows the fork()ed writer would store: length plist, photons[..][:plist]
    off = vec['list_offsets']
    closed = np.array([min(int(((np.asarray(s) >> np.uint64(56)) == 255).sum()), secs) for s in streams])
    npt = len(streams) * npix
    for s in range(secs):
        for r in range(len(streams)):
            if s >= closed[r]:
                continue                                 # second never closed: the reference writes no row
            for p in range(npix):
                k = s * npt + r * npix + p
                n = int(ref['list_len'][s, r * npix + p])
                assert n == off[k + 1] - off[k]
                assert np.array_equal(ref['lists'][s, r * npix + p, :n], vec['list_words'][off[k]:off[k + 1]])
    return ref, vec


def test_reference_loop_ragged_corrupt_eos_and_cap_quirk():
    streams = _ragged_streams(21, 3, 37, 4, 30000)
    ref, vec = _compare(streams, 37, 4, 300)
    assert ref['counts'].max() == 299 and ref['n_corrupt_eos'] == 1 and ref['n_nonpixel'] > 0
    lit = odec.packetmaster_bin_literal(_ragged_streams(3, 3, 37, 2, 3000), 37, 2, 300)
    ref2 = pm_ref.run(_ragged_streams(3, 3, 37, 2, 3000), 37, 2, 300, want_lists=True)
    assert np.array_equal(lit['counts'], ref2['counts'])
    for (sec, pix), words in lit['lists'].items():
        n = int(ref2['list_len'][sec, pix])
        assert ref2['lists'][sec, pix, :n].tolist() == words


def test_reference_loop_dense_eos_and_empty_stream():
    rng = np.random.default_rng(2)
    st0 = odec.pack_word(rng.integers(0, 5, 20000), 1, 2, 3, 4)
    st0[rng.random(st0.size) < 0.15] = np.uint64(0xFFFFFFFFFFFFFFFF)
    streams = [st0, np.zeros(0, dtype=np.uint64)]
    ref, vec = _compare(streams, 5, 50, 2500)
    assert list(ref['sec']) == [50, 0]


def test_reference_loop_cap_2500_hot_pixels():
    """The real cap: a pixel with 3000 words in a second keeps 2499 (slot 2499 is overwritten, PacketMaster.c:373-380)."""
    rng = np.random.default_rng(5)
    streams = []
    for r in range(2):
        parts = []
        for s in range(3):
            n = 40000
            ch = rng.integers(0, 253, n)
            ch[:3000] = 7
            rng.shuffle(ch)
            parts += [odec.pack_word(ch, rng.integers(0, 4096, n), rng.integers(0, 4096, n), rng.integers(0, 4096, n),
                                     np.sort(rng.integers(0, 10 ** 6, n))), np.array([2 ** 64 - 1], dtype=np.uint64)]
        streams.append(np.concatenate(parts))
    ref, vec = _compare(streams, 253, 3, 2500)
    assert ref['counts'][:, 7].max() == 2499 and ref['counts'][:, 253 + 7].max() == 2499


def test_reference_loop_equals_c_core_on_the_array_geometry():
    """8 roaches x 253 pixels (PacketMasterR4.c:49 geometry), 2e5 words: reference loop == packetmaster_core.c == NumPy."""
    if not pm_ref.available(8, 253, 2500):
        pytest.skip('oracle/_ref not built')
    streams = _ragged_streams(9, 8, 253, 3, 6000, hot=False, corrupt=False)
    ref = pm_ref.run(streams, 253, 3, 2500)
    so = os.path.join(ROOT, 'oracle', '_build', 'libpm_core.so')
    if not os.path.exists(so):
        import subprocess
        subprocess.check_call(['make', '-s', '-C', os.path.join(ROOT, 'oracle')])
    lib = ctypes.CDLL(so)
    counts = np.zeros((3, 8 * 253), dtype=np.int32)
    st = (ctypes.c_int64 * 5)()
    for r, w in enumerate(streams):
        w = np.ascontiguousarray(w)
        assert lib.pm_core_words(w.ctypes.data_as(ctypes.c_void_p), ctypes.c_int64(w.size), r, 253, 8 * 253, 3, 2500,
                                 counts.ctypes.data_as(ctypes.c_void_p), None, None, 0, 44, None, st) == 0
    assert np.array_equal(counts, ref['counts'])
    assert (st[0], st[1], st[2]) == (ref['n_eos'], ref['n_corrupt_eos'], ref['n_nonpixel'])
    assert np.array_equal(odec.packetmaster_bin(streams, 253, 3)['counts'], ref['counts'])
