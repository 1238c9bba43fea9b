"""GPU: the full chain (K4 -> K5 -> K6) through the public ReadoutChain API."""
import numpy as np
import pytest

from oracle import channelizer as oc
from oracle import decode as odec

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def ctx():
    from mkids_sdr_b200 import _lib
    return _lib.default_context(0)


def test_chain_counts_and_hist_match_oracle(ctx):
    from mkids_sdr_b200.chain import ReadoutChain
    from mkids_sdr_b200.channelizer import synth_adc
    B, n_lut, n = 2, 2 ** 16, 2 ** 20
    chain, boards = ReadoutChain.synthetic(B, n_lut, 40, seed0=11, threshold=-2400, holdoff=100, ctx=ctx, exptime=3,
                                           n_bins=16)
    tb = np.stack([bd['tone_bins'] for bd in boards])
    iq = synth_adc(B, n, tb, n_lut=n_lut, pulse_rate=4000., seed=21, ctx=ctx)
    cap = chain.chan.words_capacity(n)
    wh = np.zeros((B, cap), dtype=np.uint64)
    streams = [[] for _ in range(B)]
    for rep in range(2):                       # two consecutive batches: streaming state is carried
        nw = chain.process(iq, words_host=wh)
        for b in range(B):
            streams[b].append(wh[b, :nw[b]].copy())
    streams = [np.concatenate(s) for s in streams]
    assert all(len(s) > 50 for s in streams)
    ref = odec.packetmaster_bin(streams, 253, 3)
    assert np.array_equal(chain.dec.counts_raw(), ref['raw_counts'])
    lut = np.arange(4096) * 16 // 4096
    assert np.array_equal(chain.dec.hist(), odec.pixel_field_hist(streams, 253, 3, 'peak', lut, 16))
    # the words themselves: oracle detection on the GPU's own phase rows (both batches in one stream)
    chain.chan.reset()
    _, ph = chain.chan.process(np.concatenate([iq, iq], axis=1), detect=False, want_phase=True)
    for b in range(B):
        cfg = oc.ChanConfig(boards[b]['bins'], boards[b]['I_dds'], boards[b]['Q_dds'], chain.fir_int,
                            thresholds=np.full(256, -2400), zero_ch=boards[b]['zero_ch'].astype(bool), M=20, L=100, W=32)
        want = oc.detect_emit(ph[b], cfg, 0, np.zeros(256, np.int64), 2 * n // 512 - 64 - cfg.M)
        assert np.array_equal(streams[b], np.array(want, dtype=np.uint64))


def test_thresholds_from_noise_and_trigger_rate(ctx):
    from mkids_sdr_b200.chain import ReadoutChain
    from mkids_sdr_b200.channelizer import synth_adc
    chain, boards = ReadoutChain.synthetic(1, 2 ** 16, 24, seed0=5, ctx=ctx, exptime=2, holdoff=1000)
    thr = chain.derive_thresholds(boards, n=2 ** 24 // 2)
    act = ~boards[0]['zero_ch'].astype(bool)
    assert (thr[0][act] < -10).all() and (thr[0][act] > -25736).all()
    n = 2 ** 23
    iq = synth_adc(1, n, boards[0]['tone_bins'][None, :], n_lut=2 ** 16, pulse_rate=1000., seed=9, ctx=ctx)
    nw = chain.process(iq)
    # 16 ms of data, 24 channels, ~1000 pulses/s (depth 20-120 deg): tens to a few hundred words
    assert 100 < nw[0] < 24 * 40, nw


def test_chain_async_equals_sync(ctx):
    """process_async (no host round trip: word counts and carried seconds stay on the device) gives the same
    per-pixel products, seconds and word counts as the synchronous process(), batch after batch."""
    from mkids_sdr_b200.chain import ReadoutChain
    from mkids_sdr_b200.channelizer import synth_adc
    B, n_lut, n = 2, 2 ** 16, 2 ** 20
    res = []
    for mode in ('sync', 'async', 'mixed'):
        chain, boards = ReadoutChain.synthetic(B, n_lut, 40, seed0=11, threshold=-2400, holdoff=100, ctx=ctx, exptime=3,
                                               n_bins=16)
        tb = np.stack([bd['tone_bins'] for bd in boards])
        iq = ctx.to_device(synth_adc(B, n, tb, n_lut=n_lut, pulse_rate=4000., seed=21, ctx=ctx))
        nw = None
        for rep in range(3):
            if mode == 'sync' or (mode == 'mixed' and rep == 1):
                nw = chain.process(iq, n=n)
            else:
                chain.process_async(iq, n=n)
                nw = chain.sync_state() if rep == 2 or mode == 'mixed' else None
        if nw is None:
            nw = chain.sync_state()
        res.append((chain.dec.counts_raw(), chain.dec.hist(), chain.sec.copy(), np.asarray(nw).copy()))
    for r in res[1:]:
        for a, b in zip(res[0], r):
            assert np.array_equal(a, b)
    assert res[0][0].sum() > 100


def test_full_size_config3_properties(ctx):
    """BASELINE config 3 at full size (8 boards x 2^25 samples per batch, N_lut = 2^19), checked through
    size-independent properties: streaming invariance (one batch of 2^25 == two batches of 2^24: identical photon
    words), checksum of checksums (histogram row sums == per-pixel counts == words emitted), time order."""
    from mkids_sdr_b200.chain import ReadoutChain
    from mkids_sdr_b200.channelizer import synth_adc
    B, n_lut, n = 8, 2 ** 19, 2 ** 25
    chain, boards = ReadoutChain.synthetic(B, n_lut, 253, seed0=42, ctx=ctx, exptime=4, n_bins=64)
    tb = np.stack([bd['tone_bins'] for bd in boards])
    iq = ctx.alloc(B * n * 4)
    synth_adc(B, n, tb, n_lut=n_lut, pulse_rate=1000.0, seed=1000, out=iq, ctx=ctx)
    cap = chain.chan.words_capacity(n)
    wh = np.zeros((B, cap), dtype=np.uint64)
    nw = chain.process(iq, n=n, words_host=wh)
    whole = [wh[b, :nw[b]].copy() for b in range(B)]
    counts, hist = chain.dec.counts_raw(), chain.dec.hist()
    assert nw.min() > 5000                                     # ~1000 pulses/s x 253 channels x 65.5 ms
    assert int(counts.sum()) == int(nw.sum())                  # no second boundary inside 65.5 ms: no EOS words
    assert np.array_equal(hist.sum(axis=1), counts.sum(axis=0))
    for b in range(B):
        ts = (whole[b] & np.uint64(0xFFFFF)).astype(np.int64)
        assert np.all(np.diff(ts) >= 0)                        # (time, channel) order
        assert int((whole[b] >> np.uint64(56)).max()) < 253
    # the same stream in two halves: [B][n] rows are contiguous per board, so copy each half into its own buffer
    chain.reset()
    half = ctx.alloc(B * (n // 2) * 4)
    parts = [[] for _ in range(B)]
    for h in range(2):
        for b in range(B):
            ctx._check(ctx.lib.mkid_memcpy(ctx.h, half.ptr + b * (n // 2) * 4, iq.ptr + (b * n + h * (n // 2)) * 4, (n // 2) * 4))
        nwh = chain.process(half, n=n // 2, words_host=wh)
        for b in range(B):
            parts[b].append(wh[b, :nwh[b]].copy())
    for b in range(B):
        assert np.array_equal(np.concatenate(parts[b]), whole[b])
    assert np.array_equal(chain.dec.counts_raw(), counts) and np.array_equal(chain.dec.hist(), hist)
    iq.free(); half.free()


def test_process_stream_equals_process(ctx):
    """The pipelined host-batch path (double-buffered uploads on a second stream) gives the same words, counts and
    histogram as process() batch by batch."""
    from mkids_sdr_b200.chain import ReadoutChain
    from mkids_sdr_b200.channelizer import synth_adc
    B, n_lut, n = 2, 2 ** 16, 2 ** 19
    res = []
    for mode in ('sync', 'stream'):
        chain, boards = ReadoutChain.synthetic(B, n_lut, 40, seed0=11, threshold=-2400, holdoff=100, ctx=ctx, exptime=3,
                                               n_bins=16)
        tb = np.stack([bd['tone_bins'] for bd in boards])
        batches = []
        for k in range(4):
            pb = ctx.pinned(B * n * 4)
            a = pb.view(np.int16).reshape(B, n, 2)
            a[:] = synth_adc(B, n, tb, n_lut=n_lut, pulse_rate=4000., seed=30 + k, ctx=ctx)
            batches.append((pb, a))
        cap = chain.chan.words_capacity(n)
        wh = np.zeros((B, cap), dtype=np.uint64)
        got = []
        if mode == 'sync':
            for pb, a in batches:
                nw = chain.process(a, n=n, words_host=wh)
                got.append([wh[b, :nw[b]].copy() for b in range(B)])
        else:
            for nw in chain.process_stream((a for pb, a in batches), n, words_host=wh):
                got.append([wh[b, :nw[b]].copy() for b in range(B)])
        res.append((got, chain.dec.counts_raw(), chain.dec.hist()))
    for k in range(4):
        for b in range(B):
            assert np.array_equal(res[0][0][k][b], res[1][0][k][b])
    assert np.array_equal(res[0][1], res[1][1]) and np.array_equal(res[0][2], res[1][2])
    assert sum(len(w) for w in res[0][0][0]) > 20


def test_adc_pack12_roundtrip_and_layout(ctx):
    """12-bit packed ADC stream: the GPU expansion equals the oracle's bit layout, the GPU packer is its inverse, values
    outside 12 bits are clipped and counted; ragged tail (n not a multiple of the 1024-sample tile)."""
    rng = np.random.default_rng(5)
    for n in (4, 1020, 4096 + 8, 3 * 1024 * 40 + 12):
        iq = rng.integers(-2048, 2048, size=(n, 2)).astype(np.int16)
        iq[0] = (-2048, 2047)
        iq[-1] = (2047, -2048)
        want = oc.adc_pack12(iq)
        assert np.array_equal(oc.adc_unpack12(want), iq)
        d_iq, d_pk = ctx.to_device(iq), ctx.alloc(3 * n)
        assert ctx.adc_pack12(d_iq, n, d_pk) == 0
        assert np.array_equal(d_pk.download(np.uint8, 3 * n), want)
        d_out = ctx.alloc(4 * n)
        ctx.adc_unpack12(ctx.to_device(want), n, d_out)
        assert np.array_equal(d_out.download(np.int16, 2 * n).reshape(n, 2), iq)
    big = np.array([[3000, -5], [7, -2049], [1, 1], [0, 0]], dtype=np.int16)
    d_pk = ctx.alloc(12)
    assert ctx.adc_pack12(ctx.to_device(big), 4, d_pk) == 2
    d_out = ctx.alloc(16)
    ctx.adc_unpack12(d_pk, 4, d_out)
    assert np.array_equal(d_out.download(np.int16, 8).reshape(4, 2), np.clip(big, -2048, 2047))
    from mkids_sdr_b200 import _lib
    with pytest.raises(_lib.MkidError):
        ctx.adc_unpack12(d_pk, 6, d_out)                 # not a multiple of 4 samples


def test_process_stream_packed12_equals_int16(ctx):
    """process_stream over the 12-bit packed host format (3 bytes per sample over PCIe, expanded on the GPU) gives the
    words, counts and histogram of the int16 format bit for bit (the synthetic ADC is clipped to 12 bits)."""
    from mkids_sdr_b200.chain import ReadoutChain
    from mkids_sdr_b200.channelizer import synth_adc
    B, n_lut, n = 2, 2 ** 16, 2 ** 19
    res = []
    for fmt in ('i16', 'p12'):
        chain, boards = ReadoutChain.synthetic(B, n_lut, 40, seed0=11, threshold=-2400, holdoff=100, ctx=ctx, exptime=3,
                                               n_bins=16)
        tb = np.stack([bd['tone_bins'] for bd in boards])
        batches, keep = [], []
        for k in range(3):
            a = synth_adc(B, n, tb, n_lut=n_lut, pulse_rate=4000., seed=30 + k, ctx=ctx)
            pb = ctx.pinned(B * n * (3 if fmt == 'p12' else 4))
            if fmt == 'p12':
                v = pb.view(np.uint8).reshape(B, 3 * n)
                v[:] = oc.adc_pack12(a)
            else:
                v = pb.view(np.int16).reshape(B, n, 2)
                v[:] = a
            batches.append(v); keep.append(pb)
        cap = chain.chan.words_capacity(n)
        wh = np.zeros((B, cap), dtype=np.uint64)
        got = []
        for nw in chain.process_stream(iter(batches), n, words_host=wh, adc_format=fmt):
            got.append([wh[b, :nw[b]].copy() for b in range(B)])
        res.append((got, chain.dec.counts_raw(), chain.dec.hist()))
    for k in range(3):
        for b in range(B):
            assert np.array_equal(res[0][0][k][b], res[1][0][k][b])
    assert np.array_equal(res[0][1], res[1][1]) and np.array_equal(res[0][2], res[1][2])
    assert sum(len(w) for w in res[0][0][0]) > 20


def test_stress_config_shapes(ctx):
    """BASELINE config 4 (20 000 resonators over 10 feedlines = 80 board streams of 250 channels, per-pixel 4096-bin
    histograms [20 000][4096]): one GPU's share of an 8-GPU run (10 boards, roach ids 30..39) at a short batch, checked
    through checksums: histogram row sums == per-pixel counts == words emitted, nothing outside this rank's pixels."""
    from mkids_sdr_b200.chain import ReadoutChain
    from mkids_sdr_b200.channelizer import synth_adc
    B, n_lut, n, R_total, npix = 10, 2 ** 16, 2 ** 21, 80, 250
    chain, boards = ReadoutChain.synthetic(B, n_lut, npix, seed0=130, threshold=-2600, ctx=ctx, exptime=2, npix_per_roach=npix,
                                           n_roaches_total=R_total, roach0=30, hist_field='peak', n_bins=4096)
    tb = np.stack([bd['tone_bins'] for bd in boards])
    iq = ctx.alloc(B * n * 4)
    synth_adc(B, n, tb, n_lut=n_lut, pulse_rate=3000.0, seed=77, out=iq, ctx=ctx)
    for _ in range(2):
        chain.process_async(iq, n=n)
    nw = chain.sync_state()
    counts, hist = chain.dec.counts_raw(), chain.dec.hist()
    assert hist.shape == (R_total * npix, 4096) and counts.shape == (2, R_total * npix)
    per_pix = counts.sum(axis=0)
    assert np.array_equal(hist.sum(axis=1), per_pix)
    assert per_pix[:30 * npix].sum() == 0 and per_pix[40 * npix:].sum() == 0
    assert per_pix[30 * npix:40 * npix].sum() > 1000
    iq.free()


def test_pipelined_chain_equals_the_single_stream_chain(ctx):
    """ReadoutChain(pipelined=True): detection / decode / merged list of batch k on a second context under the channelizer
    kernel of batch k + 1 (mkid_chan_process(detect=2) + mkid_chan_detect_pending + mkid_stream_wait_event).  Same photon
    words, per-pixel products and merged lists as the single-stream chain, batch by batch."""
    from mkids_sdr_b200 import _lib
    from mkids_sdr_b200.chain import ReadoutChain
    from mkids_sdr_b200.channelizer import synth_adc
    n, nb = 2 ** 19, 5
    outs = []
    for pipelined in (False, True):
        chain, boards = ReadoutChain.synthetic(2, 2 ** 16, 40, seed0=11, threshold=-2500, holdoff=100, ctx=ctx, exptime=4,
                                               n_bins=4096, want_merged=True, pipelined=pipelined)
        tb = np.stack([bd['tone_bins'] for bd in boards])
        iq = ctx.alloc(2 * nb * n * 4)
        synth_adc(2, nb * n, tb, n_lut=2 ** 16, pulse_rate=4000., seed=21, out=iq, ctx=ctx)
        full = iq.download(np.int16).reshape(2, nb * n, 2)
        batches = [ctx.to_device(np.ascontiguousarray(full[:, k * n:(k + 1) * n])) for k in range(nb)]
        per_batch = []
        for k in range(nb):
            chain.process_async(batches[k], n=n)
            if k in (1, nb - 1):                       # look at two of the batches (a sync in between is allowed)
                nw = chain.sync_state()
                words = chain._words_dev.download(np.uint64).reshape(2, -1)
                offs = chain.merged_offsets_dev.download(np.int32, _lib.MERGE_MAX_SEC * 2 + 1)
                merged = chain.merged_words_dev.download(np.uint64, int(offs[-1]))
                per_batch.append((nw.copy(), [words[b, :nw[b]].copy() for b in range(2)], offs, merged))
        outs.append((per_batch, chain.dec.counts_raw().copy(), chain.dec.hist().copy()))
        for bf in batches:
            bf.free()
        iq.free()
    (pa, ca, ha), (pb, cb, hb) = outs
    assert ca.sum() > 300 and np.array_equal(ca, cb) and np.array_equal(ha, hb)
    for (nwa, wa, oa, ma), (nwb, wb, ob, mb) in zip(pa, pb):
        assert np.array_equal(nwa, nwb) and np.array_equal(oa, ob) and np.array_equal(ma, mb)
        assert all(np.array_equal(x, y) for x, y in zip(wa, wb))
