"""GPU parity: K6 decode / binning / histogram through the C ABI vs the oracle (bit-exact)."""
import os

import numpy as np
import pytest

from oracle import decode as odec
from oracle import pm_ref

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def ctx():
    from mkids_sdr_b200 import _lib
    return _lib.default_context(0)


def _check(dec, streams, npix, exptime, cap, field='peak', bin_lut=None, n_bins=4096):
    ref = odec.packetmaster_bin(streams, npix, exptime, cap)
    assert np.array_equal(dec.counts_raw(), ref['raw_counts'])
    assert np.array_equal(dec.counts(), ref['counts'])
    st = dec.stats_dict()
    for k in ('n_eos', 'n_corrupt_eos', 'n_nonpixel', 'n_ignored'):
        assert st[k] == ref[k], k
    if field:
        h = odec.pixel_field_hist(streams, npix, exptime, field, bin_lut, n_bins)
        assert np.array_equal(dec.hist(), h)
    # ... and against the reference's own loop (PacketMaster.c:304-397 compiled by oracle/build_pm_ref.py)
    if pm_ref.available(len(streams), npix, cap):
        r = pm_ref.run(streams, npix, exptime, cap)
        assert np.array_equal(dec.counts(), r['counts'])
        for k in ('n_eos', 'n_corrupt_eos', 'n_nonpixel'):
            assert st[k] == r[k], k


def _ragged_streams(seed, R, npix, secs, per_sec, hot=True):
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
            eos = np.array([0xFFFFFFFFFFFFFFFF if (s != 1 or r != 0) else 0xFF00000000000001], dtype=np.uint64)
            parts += [w, eos]
        streams.append(np.concatenate(parts))
    return streams


@pytest.mark.parametrize('field,n_bins', [('peak', 4096), ('base', 4096), ('p1', 10), (None, 0)])
def test_decode_words_ragged(ctx, field, n_bins):
    from mkids_sdr_b200.decode import PhotonDecoder
    R, npix, secs, cap = 3, 37, 4, 300
    streams = _ragged_streams(21, R, npix, secs, 30000)
    lut = None
    if field == 'p1':
        lut = (np.arange(4096) * 11 // 4096)          # 11 classes, class 10 is out of range for n_bins=10
    dec = PhotonDecoder(R, npix, secs, cap, field, n_bins, lut, ctx=ctx)
    dec.feed_streams(streams)
    _check(dec, streams, npix, secs, cap, field, lut, n_bins)
    assert dec.counts().max() == cap - 1
    assert list(dec.sec) == [secs + 1] * R


def test_decode_many_eos_and_empty(ctx, n_bins=16):
    from mkids_sdr_b200.decode import PhotonDecoder
    R, npix, secs, cap = 2, 5, 50, 2500
    rng = np.random.default_rng(2)
    # dense EOS: a second is closed every ~7 words, several per 2-word load group
    st0 = odec.pack_word(rng.integers(0, npix, 20000), 1, 2, 3, 4)
    st0[rng.random(st0.size) < 0.15] = np.uint64(0xFFFFFFFFFFFFFFFF)
    st1 = np.zeros(0, dtype=np.uint64)                # empty stream
    streams = [st0, st1]
    dec = PhotonDecoder(R, npix, secs, cap, 'peak', n_bins, ctx=ctx)
    dec.feed_streams(streams)
    _check(dec, streams, npix, secs, cap, 'peak', None, n_bins)


def test_decode_piecewise_equals_whole(ctx):
    """Feeding a stream in arbitrary (odd-sized, unaligned) pieces with carried seconds gives the same sums."""
    from mkids_sdr_b200.decode import PhotonDecoder
    R, npix, secs, cap = 2, 253, 6, 2500
    streams = _ragged_streams(5, R, npix, secs, 20000, hot=False)
    dec = PhotonDecoder(R, npix, secs, cap, 'peak', 4096, ctx=ctx)
    rng = np.random.default_rng(0)
    for r, st in enumerate(streams):
        cuts = np.sort(rng.choice(np.arange(1, st.size), 9, replace=False))
        pieces = np.split(st, cuts)
        sec = 0
        for pc in pieces:
            words = np.concatenate([np.zeros(1, np.uint64), pc])      # force odd alignment of the segment
            out = dec.decode_words(words, [1, 1 + pc.size], [r], [sec], n_words=words.size)
            sec = int(out[0])
    _check(dec, streams, npix, secs, cap)


@pytest.mark.parametrize('scale', [1, 40])
def test_decode_words_dev_segments(ctx, scale):
    """The device-chained entry point (segment lengths and carried seconds in device memory): short segments take the
    one-launch path (decode_small_kernel), long ones the range machinery; dense and corrupt end-of-second words,
    non-pixel channels, capacity > length, seconds beyond exptime, two calls with carried seconds."""
    from mkids_sdr_b200.decode import PhotonDecoder
    R, npix, secs, cap_ev = 3, 37, 6, 2500
    rng = np.random.default_rng(17)
    streams = []
    for r in range(R):
        n = (3000 + 500 * r) * scale
        st = odec.pack_word(rng.integers(0, npix + 3, n), rng.integers(0, 4096, n), rng.integers(0, 4096, n),
                            rng.integers(0, 4096, n), np.sort(rng.integers(0, 10 ** 6, n)))
        eos = rng.random(n) < (0.002 / scale)
        st[eos] = np.uint64(0xFFFFFFFFFFFFFFFF)
        st[np.flatnonzero(eos)[1::5]] = np.uint64(0xFF00000000000123)          # corrupt end-of-second words
        streams.append(st)
    dec = PhotonDecoder(R, npix, secs, cap_ev, 'base', 4096, ctx=ctx)
    seg_cap = np.array([len(s) + 777 for s in streams], dtype=np.int64)      # capacity > length
    seg_start = np.concatenate([[0], np.cumsum(seg_cap)[:-1]]).astype(np.int64)
    sec_a, sec_b = ctx.to_device(np.zeros(R, np.int32)), ctx.to_device(np.zeros(R, np.int32))
    halves = [[s[:len(s) // 3] for s in streams], [s[len(s) // 3:] for s in streams]]
    for part, (sin, sout) in zip(halves, ((sec_a, sec_b), (sec_b, sec_a))):
        buf = np.full(int(seg_cap.sum()), 0x1234567812345678, dtype=np.uint64)   # junk beyond the lengths
        for r in range(R):
            buf[seg_start[r]:seg_start[r] + len(part[r])] = part[r]
        dw = ctx.to_device(buf)
        ln = ctx.to_device(np.array([len(x) for x in part], dtype=np.int32))
        dec.decode_words_dev(dw, seg_start, seg_cap, ln, np.arange(R), sin, sout, buf.size)
        ctx.sync()
        dw.free(); ln.free()
    ref = odec.packetmaster_bin(streams, npix, secs, cap_ev)
    assert np.array_equal(dec.counts_raw(), ref['raw_counts'])
    assert np.array_equal(dec.hist(), odec.pixel_field_hist(streams, npix, secs, 'base', None, 4096))
    sec_out = sec_a.download(np.int32, R)
    assert list(sec_out) == [int(((s >> np.uint64(56)) == 255).sum()) for s in streams]
    assert ref['n_ignored'] > 0 and ref['n_corrupt_eos'] > 0


@pytest.mark.parametrize('which', ['ragged_peak', 'ragged_base', 'ragged_lut', 'dense_eos', 'piecewise', 'dev_segments',
                                   'wire', 'uniform_253x4096'])
def test_partitioned_histogram_forms(ctx, monkeypatch, which):
    """The histogram built from partitioned 16-bit keys (decode.cu HIST == 3, the MKID_DEC_PART=1 form of a histogram too
    large for shared memory) forced on the inputs of the other tests: hot pixel whose tile overflows its bucket (the surplus
    goes to the histogram directly), low-half field, bin LUT with out-of-range classes, dense / corrupt end-of-second words
    and an empty stream, unaligned pieces with carried seconds, device-resident segment lengths, seconds beyond exptime
    (taken back by the commit kernel), the wire format; and that MKID_DEC_PART=0 gives the same on a plain input."""
    from mkids_sdr_b200 import synth
    from mkids_sdr_b200.decode import PhotonDecoder
    monkeypatch.setenv('MKID_DEC_PART', '1')
    if which == 'ragged_peak':
        test_decode_words_ragged(ctx, 'peak', 4096)
    elif which == 'ragged_base':
        test_decode_words_ragged(ctx, 'base', 4096)
    elif which == 'ragged_lut':
        R, npix, secs, cap = 3, 37, 4, 300
        streams = _ragged_streams(22, R, npix, secs, 30000)
        lut = np.arange(4096) * 311 // 4096                       # 311 classes, those >= 300 are out of range
        dec = PhotonDecoder(R, npix, secs, cap, 'p1', 300, lut, ctx=ctx)
        dec.feed_streams(streams)
        _check(dec, streams, npix, secs, cap, 'p1', lut, 300)
    elif which == 'dense_eos':
        test_decode_many_eos_and_empty(ctx, 1024)
    elif which == 'piecewise':
        test_decode_piecewise_equals_whole(ctx)
    elif which == 'dev_segments':
        test_decode_words_dev_segments(ctx, 40)
    elif which == 'wire':
        test_decode_wire_bundles(ctx)
    else:
        R, npix, secs = 8, 253, 3
        streams, _ = synth.photon_streams(4 * 10 ** 6, R, npix, secs, seed=31, n_hot=3, hot_rate=3000)
        hists = []
        for mode in ('1', '0'):
            monkeypatch.setenv('MKID_DEC_PART', mode)
            dec = PhotonDecoder(R, npix, secs, 2500, 'peak', 4096, ctx=ctx)
            dec.feed_streams(streams)
            hists.append(dec.hist())
            assert np.array_equal(dec.counts_raw(), odec.packetmaster_bin(streams, npix, secs)['raw_counts'])
        assert np.array_equal(hists[0], hists[1])
        assert np.array_equal(hists[0], odec.pixel_field_hist(streams, npix, secs, 'peak'))
        assert int(hists[0].sum()) > 3 * 10 ** 6


def test_quicklook_matches_reference_run(ctx, golden_dir):
    """decode.QuickLook on the GPU against pulses.QuickLook itself (lib/pulses.py:210-236, run by
    tests/golden/make_golden_refrun.py on a stand-in observation file): the decoder's count array is loaded with the
    golden per-second counts, the preview must be the image the reference handed to imshow."""
    from mkids_sdr_b200 import _lib
    from mkids_sdr_b200.decode import PhotonDecoder, QuickLook
    g = np.load(os.path.join(golden_dir, 'quicklook_golden.npz'))
    counts = np.ascontiguousarray(g['ql_counts'], dtype=np.uint32)
    secs, n_pix = counts.shape
    dec = PhotonDecoder(8, n_pix // 8, secs, 2500, None, 1, ctx=ctx)
    ctx._check(ctx.lib.mkid_memcpy(ctx.h, _lib.ptr(dec.counts_dev), _lib.ptr(counts), counts.nbytes))
    ctx.sync()
    for i in range(3):
        t0, t1 = (int(v) for v in g['ql_span_%d' % i])
        assert np.array_equal(QuickLook(dec, g['ql_adr'], t0, t1), g['ql_skysub_%d' % i]), (t0, t1)


def test_decode_wire_bundles(ctx):
    from mkids_sdr_b200 import synth
    from mkids_sdr_b200.decode import PhotonDecoder
    R, npix, secs, cap = 4, 253, 3, 2500
    streams, _ = synth.photon_streams(400000, R, npix, secs, seed=77, n_hot=2, hot_rate=3000)
    wire = synth.streams_to_wire(streams)
    dec = PhotonDecoder(R, npix, secs, cap, 'peak', 4096, ctx=ctx)
    for r in range(R):           # bundle by bundle, as PacketMaster receives them (interleaved roaches)
        pass
    nb = [w.size // 65536 for w in wire]
    for b in range(max(nb)):
        for r in range(R):
            if b < nb[r]:
                dec.feed_bundles(r, wire[r][b * 65536:(b + 1) * 65536])
    _check(dec, streams, npix, secs, cap)
    assert dec.counts().max() == cap - 1            # hot pixels exceed the 2500 cap
    # all bundles in one call
    dec2 = PhotonDecoder(R, npix, secs, cap, 'peak', 4096, ctx=ctx)
    allw = np.concatenate(wire)
    off = np.concatenate([[0], np.cumsum(nb)])
    dec2.decode_wire(allw, off, np.arange(R))
    _check(dec2, streams, npix, secs, cap)
    # wire -> words oracle agrees with the synth packer
    assert np.array_equal(odec.wire_to_words(wire[0].tobytes()), streams[0])


def test_quicklook_and_unpack_and_reinterpret(ctx, golden_dir):
    import os
    from mkids_sdr_b200 import synth
    from mkids_sdr_b200.decode import PhotonDecoder, unpack_fields
    from mkids_sdr_b200.Utils import binTools
    R, npix, secs = 8, 253, 2
    streams, _ = synth.photon_streams(300000, R, npix, secs, seed=3)
    dec = PhotonDecoder(R, npix, secs, ctx=ctx)
    dec.feed_streams(streams)
    rng = np.random.default_rng(1)
    pixel_adr = rng.permutation(R * npix).reshape(46, 44)
    img = dec.quicklook_image(1, pixel_adr)
    assert np.array_equal(img, odec.quicklook_image(dec.counts()[1], pixel_adr))
    f = unpack_fields(streams[2])
    for a, b in zip(f, odec.unpack_fields(streams[2])):
        assert np.array_equal(a, b)
    g = np.load(os.path.join(golden_dir, 'utils_bin_py3.npz'))
    for nb, bp, key in ((12, 9, 'reinterpret_12_9'), (16, 13, 'reinterpret_16_13'), (18, 16, 'reinterpret_18_16')):
        assert np.array_equal(binTools.reinterpretBin(g['values'], nb, bp), g[key])


def test_full_size_config1_properties(ctx):
    """BASELINE config 0 at full size (1e7 words, 2024 pixels): checked through size-independent
    properties and against the oracle's vectorised counts."""
    from mkids_sdr_b200 import synth
    from mkids_sdr_b200.decode import PhotonDecoder
    R, npix, secs = 8, 253, 10
    streams, eos = synth.photon_streams(10 ** 7, R, npix, secs, seed=1234)
    dec = PhotonDecoder(R, npix, secs, 2500, 'peak', 4096, ctx=ctx)
    dec.feed_streams(streams)
    raw = dec.counts_raw()
    st = dec.stats_dict()
    assert st['n_eos'] == R * secs and st['n_corrupt_eos'] == 0
    assert int(raw.sum()) == st['n_valid'] == 10 ** 7
    assert int(dec.hist().sum()) == 10 ** 7
    assert np.array_equal(dec.hist().sum(axis=1), raw.sum(axis=0))       # checksum of checksums
    assert dec.counts().max() == 2499
    ref = odec.packetmaster_bin(streams, npix, secs)
    assert np.array_equal(raw, ref['raw_counts'])
    assert st['n_nonpixel'] == ref['n_nonpixel']


def test_photon_lists_match_packetmaster(ctx):
    """Per-(second, pixel) photon lists (PacketMaster.c:371-380): arrival order inside a key, cap quirk."""
    from mkids_sdr_b200 import synth
    from mkids_sdr_b200.decode import PhotonDecoder
    R, npix, secs, cap = 4, 253, 3, 2500
    streams, _ = synth.photon_streams(600000, R, npix, secs, seed=5, n_hot=3, hot_rate=3000)
    dec = PhotonDecoder(R, npix, secs, cap, None, 0, ctx=ctx)
    lens = [len(s) for s in streams]
    off = np.concatenate([[0], np.cumsum(lens)])
    lw, lo, sec_out = dec.decode_lists(np.concatenate(streams), off, np.arange(R))
    ref = odec.packetmaster_bin(streams, npix, secs, cap, want_lists=True)
    assert np.array_equal(lo, ref['list_offsets'])
    assert np.array_equal(lw, ref['list_words'])
    assert np.array_equal(dec.counts_raw(), ref['raw_counts'])
    assert np.diff(lo).max() == cap - 1                       # hot pixels hit the cap
    # literal word-by-word loop on a small ragged case, two segments per roach (carried seconds)
    R2, npix2, secs2, cap2 = 2, 11, 4, 40
    small = _ragged_streams(3, R2, npix2, secs2, 900)
    dec2 = PhotonDecoder(R2, npix2, secs2, cap2, None, 0, ctx=ctx)
    lw2, lo2, _ = dec2.decode_lists(np.concatenate(small), np.concatenate([[0], np.cumsum([len(s) for s in small])]),
                                    np.arange(R2))
    lit = odec.packetmaster_bin_literal(small, npix2, secs2, cap2)
    for s in range(secs2):
        for pix in range(R2 * npix2):
            k = s * R2 * npix2 + pix
            got = lw2[lo2[k]:lo2[k + 1]].tolist()
            want = [int(w) for w in lit['lists'].get((s, pix), [])]
            assert got == want, (s, pix)


def test_merged_photon_list(ctx):
    """Time-ordered merged list of config 4: (second, roach) keys, stream order inside, no cap; literal check of the
    ordering on a small ragged case fed as two segments per roach with carried seconds."""
    from mkids_sdr_b200 import synth
    from mkids_sdr_b200.decode import PhotonDecoder
    R, npix, secs = 8, 253, 5
    streams, _ = synth.photon_streams(2 * 10 ** 6, R, npix, secs, seed=11, n_hot=3, hot_rate=3000)
    dec = PhotonDecoder(R, npix, secs, 2500, None, 0, ctx=ctx)
    off = np.concatenate([[0], np.cumsum([len(s) for s in streams])])
    lw, lo, sec_out = dec.decode_merged(np.concatenate(streams), off, np.arange(R))
    ref_w, ref_o = odec.merged_list(streams, npix, secs)
    assert np.array_equal(lo, ref_o)
    assert np.array_equal(lw, ref_w)
    assert np.array_equal(dec.counts_raw(), odec.packetmaster_bin(streams, npix, secs)['raw_counts'])
    ts = (lw & np.uint64(0xFFFFF)).astype(np.int64)                 # timestamps only grow inside a (second, roach)
    for k in range(secs * R):
        assert np.all(np.diff(ts[lo[k]:lo[k + 1]]) >= 0)
    # ragged, more seconds in the stream than exptime, second segment of every roach continues the first
    R2, npix2, secs2 = 3, 37, 4
    small = _ragged_streams(9, R2, npix2, secs2 + 1, 5000)
    halves, roach, seg_sec = [], [], []
    for r, st in enumerate(small):
        cut = st.size // 2 + r
        n_eos_first = int(((st[:cut] >> np.uint64(56)) == 255).sum())
        halves += [st[:cut], st[cut:]]
        roach += [r, r]
        seg_sec += [0, n_eos_first]
    dec2 = PhotonDecoder(R2, npix2, secs2, 2500, None, 0, ctx=ctx)
    off2 = np.concatenate([[0], np.cumsum([h.size for h in halves])])
    lw2, lo2, so2 = dec2.decode_merged(np.concatenate(halves), off2, roach, seg_sec)
    ref_w2, ref_o2 = odec.merged_list(small, npix2, secs2)
    assert np.array_equal(lo2, ref_o2) and np.array_equal(lw2, ref_w2)
    # segments out of time order (second half of every roach first): a key that straddles the cut is revisited
    firsts, seconds = halves[0::2], halves[1::2]
    rev = [h for pair in zip(seconds, firsts) for h in pair]
    rev_sec = [v for pair in zip(seg_sec[1::2], seg_sec[0::2]) for v in pair]
    off4 = np.concatenate([[0], np.cumsum([h.size for h in rev])])
    wa, oa = odec.merged_list(seconds, npix2, secs2, sec0=seg_sec[1::2])
    wb, ob = odec.merged_list(firsts, npix2, secs2)
    want = np.concatenate([np.concatenate([wa[oa[k]:oa[k + 1]], wb[ob[k]:ob[k + 1]]]) for k in range(secs2 * R2)])
    dec4 = PhotonDecoder(R2, npix2, secs2, 2500, None, 0, ctx=ctx)
    lw4, lo4, _ = dec4.decode_merged(np.concatenate(rev), off4, roach, rev_sec)
    assert np.array_equal(lo4, ref_o2) and np.array_equal(lw4, want)
    # the same input through the per-pixel lists (cap far away): every key holds its second-half words first
    dec5 = PhotonDecoder(R2, npix2, secs2, 10 ** 6, None, 0, ctx=ctx)
    lw5, lo5, _ = dec5.decode_lists(np.concatenate(rev), off4, roach, rev_sec)
    la = odec.packetmaster_bin([np.concatenate([np.full(s0, 0xFFFFFFFFFFFFFFFF, np.uint64), h]) for s0, h in zip(seg_sec[1::2], seconds)],
                               npix2, secs2, 10 ** 6, want_lists=True)
    lb = odec.packetmaster_bin(firsts, npix2, secs2, 10 ** 6, want_lists=True)
    wa5, oa5, wb5, ob5 = la['list_words'], la['list_offsets'], lb['list_words'], lb['list_offsets']
    want5 = np.concatenate([np.concatenate([wa5[oa5[k]:oa5[k + 1]], wb5[ob5[k]:ob5[k + 1]]]) for k in range(secs2 * R2 * npix2)])
    assert np.array_equal(lw5, want5)
    assert np.array_equal(np.diff(lo5), np.diff(oa5) + np.diff(ob5))
    # empty input
    dec3 = PhotonDecoder(R2, npix2, secs2, 2500, None, 0, ctx=ctx)
    lw3, lo3, _ = dec3.decode_merged(np.zeros(0, np.uint64), [0, 0], [1])
    assert lw3.size == 0 and not lo3.any()


def test_lists_long_ranges(ctx):
    """Ranges of several 2048-word sort blocks (the range table only grows beyond one block above ~1e7 words): lists
    and merged list of 2.4e7 words against the oracle."""
    from mkids_sdr_b200 import synth
    from mkids_sdr_b200.decode import PhotonDecoder
    R, npix, secs, cap = 8, 253, 10, 2500
    streams, _ = synth.photon_streams(24 * 10 ** 6, R, npix, secs, seed=77, n_hot=5, hot_rate=3000)
    off = np.concatenate([[0], np.cumsum([len(s) for s in streams])])
    words = ctx.to_device(np.concatenate(streams))
    ref = odec.packetmaster_bin(streams, npix, secs, cap, want_lists=True)
    dec = PhotonDecoder(R, npix, secs, cap, None, 0, ctx=ctx)
    lw, lo, _ = dec.decode_lists(words, off, np.arange(R), n_words=int(off[-1]))
    assert np.array_equal(lo, ref['list_offsets'])
    assert np.array_equal(lw, ref['list_words'])
    dec2 = PhotonDecoder(R, npix, secs, cap, None, 0, ctx=ctx)
    mw, mo, _ = dec2.decode_merged(words, off, np.arange(R), n_words=int(off[-1]))
    ref_w, ref_o = odec.merged_list(streams, npix, secs)
    assert np.array_equal(mo, ref_o) and np.array_equal(mw, ref_w)
    words.free()


def test_lists_from_wire_bundles(ctx):
    """Per-pixel lists and the merged list straight from PulseServer bundles equal those of the flat streams (the
    channel-254 filler that pads the last bundle is a non-pixel word)."""
    from mkids_sdr_b200 import synth
    from mkids_sdr_b200.decode import PhotonDecoder
    R, npix, secs, cap = 4, 253, 3, 2500
    streams, _ = synth.photon_streams(500000, R, npix, secs, seed=21, n_hot=2, hot_rate=3000)
    wire = synth.streams_to_wire(streams)
    nb = [w.size // 65536 for w in wire]
    off = np.concatenate([[0], np.cumsum(nb)])
    allw = np.concatenate(wire)
    ref = odec.packetmaster_bin(streams, npix, secs, cap, want_lists=True)
    dec = PhotonDecoder(R, npix, secs, cap, None, 0, ctx=ctx)
    lw, lo, _ = dec.decode_wire_lists(allw, off, np.arange(R))
    assert np.array_equal(lo, ref['list_offsets']) and np.array_equal(lw, ref['list_words'])
    assert np.array_equal(dec.counts_raw(), ref['raw_counts'])
    dec2 = PhotonDecoder(R, npix, secs, cap, None, 0, ctx=ctx)
    mw, mo, _ = dec2.decode_wire_lists(allw, off, np.arange(R), merged=True)
    ref_w, ref_o = odec.merged_list(streams, npix, secs)
    assert np.array_equal(mo, ref_o) and np.array_equal(mw, ref_w)


def test_dashboard_make_image_matches_reference_run(ctx, golden_dir):
    """decode.Dashboard (mkid_dashboard_image + the twin of make_image's bookkeeping) against the reference's own
    StartQt4.make_image (ArconsDashboard.py:633-723) executed in the dev container over 7 seconds: sky taking, sky
    subtraction (with the reference's orientation quirk), 3-second integration window that excludes the current
    second, flat field, brightest-pixel contrast and the saturated-pixel list."""
    import os
    from mkids_sdr_b200 import _lib
    from mkids_sdr_b200.decode import Dashboard, PhotonDecoder
    g = np.load(os.path.join(golden_dir, 'refrun_golden.npz'))
    counts = g['dash_counts'].astype(np.uint32)
    secs, n_pix = counts.shape
    dec = PhotonDecoder(8, 253, secs, 2500, None, 0, ctx=ctx)
    ctx._check(ctx.lib.mkid_memcpy(ctx.h, _lib.ptr(dec.counts_dev), _lib.ptr(np.ascontiguousarray(counts)), counts.nbytes))
    ctx.sync()
    db = Dashboard(dec, g['dash_pixel_adr'])
    db.taking_sky, db.skytime, db.int_time, db.brightpix = True, 2, 3, 5
    db.flatFactors = g['dash_flat']
    for t in range(secs):
        db.sky_subtraction = t >= 3
        db.flat_field = t >= 5
        frame = db.make_image()
        assert np.array_equal(frame, g['dash_frame_%d' % t]), t
        assert db.vmax == float(g['dash_vmax_%d' % t]), t
        assert np.array_equal(np.array(db.redpix), g['dash_redpix_%d' % t]), t
    assert np.array_equal(db.skyrate, g['dash_skyrate'])


def test_merge_words_dev_matches_numpy(ctx):
    """Device-chained merged list of one batch (mkid_merge_words_dev): segments with end-of-second words, non-pixel
    channels, junk beyond the device-side lengths, a carried second that runs into exptime."""
    import ctypes
    from mkids_sdr_b200 import _lib
    from mkids_sdr_b200.decode import PhotonDecoder
    R, npix, exptime = 5, 37, 6
    rng = np.random.default_rng(3)
    caps = np.array([9000, 3000, 1, 20000, 7000], dtype=np.int64)
    lens = np.array([8100, 0, 1, 17777, 6999], dtype=np.int32)
    sec0 = np.array([0, 2, 5, 4, 3], dtype=np.int32)
    start = np.concatenate([[0], np.cumsum(caps)[:-1]]).astype(np.int64)
    buf = rng.integers(0, 2 ** 63, int(caps.sum()), dtype=np.int64).astype(np.uint64)       # junk everywhere
    segs = []
    for i in range(R):
        n = int(lens[i])
        w = odec.pack_word(rng.integers(0, npix + 3, n), rng.integers(0, 4096, n), rng.integers(0, 4096, n),
                           rng.integers(0, 4096, n), np.sort(rng.integers(0, 10 ** 6, n)))
        if n > 100:
            eos = np.sort(rng.choice(np.arange(10, n), 3 if i != 3 else 6, replace=False))   # segment 3: more than MERGE_MAX_SEC - 1
            w[eos] = np.uint64(0xFFFFFFFFFFFFFFFF)
        buf[start[i]:start[i] + n] = w
        segs.append(w)
    dec = PhotonDecoder(R, npix, exptime, ctx=ctx)
    dw, dl, ds = ctx.to_device(buf), ctx.to_device(lens), ctx.to_device(sec0)
    out = ctx.alloc(int(caps.sum()) * 8)
    off = ctx.alloc((_lib.MERGE_MAX_SEC * R + 1) * 4)
    ctx._check(ctx.lib.mkid_merge_words_dev(ctx.h, _lib.ptr(dw), _lib.ptr(start), _lib.ptr(caps), _lib.ptr(dl), _lib.ptr(ds), R,
                                            ctypes.byref(dec.cfg), _lib.ptr(out), int(caps.sum()), _lib.ptr(off)))
    ctx.sync()
    offs = off.download(np.int32, _lib.MERGE_MAX_SEC * R + 1)
    words = out.download(np.uint64, int(offs[-1]))
    exp = []
    exp_off = [0]
    for ls in range(_lib.MERGE_MAX_SEC):
        for i in range(R):
            w = segs[i]
            adr = (w >> np.uint64(56)).astype(np.int64)
            is_eos = adr == 255
            ls_true = np.cumsum(is_eos) - is_eos
            key = np.minimum(ls_true, _lib.MERGE_MAX_SEC - 1)
            ok = (~is_eos) & (adr < npix) & (sec0[i] + ls_true < exptime) & (key == ls)
            exp.append(w[ok])
            exp_off.append(exp_off[-1] + int(ok.sum()))
    assert np.array_equal(offs, np.array(exp_off, dtype=np.int32))
    assert np.array_equal(words, np.concatenate(exp))
    assert offs[-1] > 20000


def test_quicklook_files_and_beammap_parse(ctx, tmp_path):
    """The per-second `<obs>_<sec>.txt` images PacketMaster leaves for the dashboard (PacketMaster.c:679-727, 1024-1045)
    and the beammap "/r%d/p%d/" parse (:880-904); the dashboard reads the file with numpy.loadtxt
    (ArconsDashboard.py:635)."""
    from mkids_sdr_b200 import synth
    from mkids_sdr_b200.decode import PhotonDecoder, QuickLookWriter, parse_beammap
    R, npix, secs = 4, 253, 3
    rows, cols = 23, 44                                                  # 1012 = 4 x 253 pixels
    rng = np.random.default_rng(8)
    perm = rng.permutation(R * npix)
    names = np.array(['/r%d/p%d/' % (a // npix, a % npix) for a in perm], dtype=object).reshape(rows, cols)
    adr = parse_beammap(names, npix)
    assert np.array_equal(adr.reshape(-1), perm)
    assert parse_beammap([['/r2/p17/t1319000000', 'r1/p5', '/rX/p3/']], npix).tolist() == [[2 * npix + 17, npix + 5, 3]]
    streams, _ = synth.photon_streams(300000, R, npix, secs, seed=4, n_hot=2, hot_rate=3000)
    dec = PhotonDecoder(R, npix, secs, ctx=ctx)
    ql = QuickLookWriter(dec, adr, str(tmp_path / 'obs_20110726-114310.h5'))
    dec.feed_streams(streams)
    # roach 0 lags behind: only the seconds EVERY roach has closed may be written
    assert ql.flush(np.array([0, 2, 2, 2])) == []
    files = ql.flush(np.array([2, 3, 3, 3]))
    assert [os.path.basename(f) for f in files] == ['obs_20110726-114310_0.txt', 'obs_20110726-114310_1.txt']
    files = ql.flush()
    assert len(files) == 3 and not any(n.startswith('lock.') for n in os.listdir(tmp_path / 'bin'))
    ref = odec.packetmaster_bin(streams, npix, secs)
    for s, f in enumerate(files):
        img = np.loadtxt(f)
        assert img.shape == (rows, cols)
        assert np.array_equal(img, odec.quicklook_image(ref['counts'][s], adr))
        assert open(f).read().split('\n')[0].endswith(' ')               # "%d " after every value
    assert ref['counts'].max() == 2499
    # pulses.QuickLook (lib/pulses.py:210-236): photons per pixel over a span of seconds, median sky taken off
    from mkids_sdr_b200.decode import QuickLook
    for t0, t1 in ((0, 3), (1, 3), (0, 1), (2, 3), (1, 1)):
        got = QuickLook(dec, adr, t0, t1)
        want = odec.quicklook_skysub(ref['counts'], adr, t0, t1)
        assert got.dtype == np.float32 and got.shape == (rows, cols) and np.array_equal(got, want), (t0, t1)
    with pytest.raises(IndexError):
        QuickLook(dec, adr, 0, secs + 1)
