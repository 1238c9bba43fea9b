"""Synthetic inputs of the shapes SURVEY.md 8(d) names (no datasets in this environment).

Photon streams follow the wire/word format of the reference (App. A.5): 64-bit words
[63:56] channel, [55:44] peak, [43:32] p1, [31:20] baseline, [19:0] timestamp (us),
one all-ones end-of-second word per roach per second (PacketMaster.c:329-333).
"""
import numpy as np

EOS = np.uint64(0xFFFFFFFFFFFFFFFF)
BUNDLE = 8192


def pack_word(ch, peak, p1, base, ts):
    ch, peak, p1, base, ts = (np.asarray(x).astype(np.uint64) for x in (ch, peak, p1, base, ts))
    return (ch << np.uint64(56)) | (peak << np.uint64(44)) | (p1 << np.uint64(32)) | (base << np.uint64(20)) | ts


def photon_streams(n_words=10 ** 7, n_roaches=8, npix_per_roach=253, n_sec=10, seed=1234, n_hot=5,
                   hot_rate=3000, pad_to_bundle=True, gaussian_heights=False):
    """Config 1 of SURVEY 8(d): returns (streams, eos_positions) with streams[r] a u64 array
    and eos_positions[r] the indices of the EOS words in streams[r]."""
    rng = np.random.default_rng(seed)
    n_pix = n_roaches * npix_per_roach
    hot = rng.choice(n_pix, size=min(n_hot, n_pix), replace=False)
    n_hot_words = len(hot) * hot_rate * n_sec
    n_bg = max(n_words - n_hot_words, 0)
    pix = rng.integers(0, n_pix, n_bg)
    sec = rng.integers(0, n_sec, n_bg)
    if len(hot):
        pix = np.concatenate([pix, np.repeat(hot, hot_rate * n_sec)])
        sec = np.concatenate([sec, np.tile(np.repeat(np.arange(n_sec), hot_rate), len(hot))])
    n = pix.size
    ts = rng.integers(0, 10 ** 6, n)
    if gaussian_heights:
        peak = np.clip(np.round(rng.normal(2048 - 600, 120, n)), 0, 4095).astype(np.int64)
        base = np.clip(np.round(rng.normal(2048, 15, n)), 0, 4095).astype(np.int64)
    else:
        peak = rng.integers(0, 4096, n)
        base = rng.integers(0, 4096, n)
    p1 = rng.integers(0, 4096, n)
    roach = pix // npix_per_roach
    adr = pix % npix_per_roach
    words = pack_word(adr, peak, p1, base, ts)
    # order: roach, second, timestamp (ts ascending within each (roach, second))
    order = np.lexsort((ts, sec, roach))
    words, roach, sec = words[order], roach[order], sec[order]
    streams, eos_pos = [], []
    key = roach * n_sec + sec
    bounds = np.searchsorted(key, np.arange(n_roaches * n_sec + 1))
    for r in range(n_roaches):
        parts, pos, at = [], [], 0
        for s in range(n_sec):
            seg = words[bounds[r * n_sec + s]:bounds[r * n_sec + s + 1]]
            parts += [seg, np.array([EOS], dtype=np.uint64)]
            at += seg.size
            pos.append(at)
            at += 1
        st = np.concatenate(parts)
        if pad_to_bundle and st.size % BUNDLE:
            fill = BUNDLE - st.size % BUNDLE          # non-pixel channel 254 filler (must be ignored)
            st = np.concatenate([st, np.full(fill, pack_word(254, 0, 0, 0, 0), dtype=np.uint64)])
        streams.append(st)
        eos_pos.append(np.array(pos, dtype=np.int64))
    return streams, eos_pos


def streams_to_wire(streams):
    """Per roach: bundles of 8192 big-endian low halves then 8192 big-endian high halves
    (PulseServer.c:338-352).  Returns a list of uint8 arrays."""
    out = []
    for st in streams:
        assert st.size % BUNDLE == 0
        w = st.reshape(-1, BUNDLE)
        buf = np.empty((w.shape[0], 2, BUNDLE), dtype='>u4')
        buf[:, 0, :] = (w & np.uint64(0xFFFFFFFF)).astype('>u4')
        buf[:, 1, :] = (w >> np.uint64(32)).astype('>u4')
        out.append(buf.reshape(-1).view(np.uint8))
    return out
