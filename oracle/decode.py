"""Oracle: photon-word decode / per-pixel binning / histogramming.

TEST INFRASTRUCTURE (see oracle/__init__.py).  NumPy restatement of
  * wire format        PulseServer.c:318-352 (32 KiB low block then 32 KiB high
                       block, big-endian u32) / PacketMaster.c:286-307
  * bitfield layout    ROACH_Pulses.py:805-811
  * binning semantics  PacketMaster.c:304-397 (EOS, non-pixel, 2500 cap quirk)
  * readPulses         ROACH_Pulses.py:782-889 (ring wrap, per-channel lists,
                       Fix12_9 -> deg, three 40-bin histograms)
  * image_Worker       ArconsDashboard.py:1331-1337 (per-pixel 10-bin spectrum)
"""
import numpy as np

BUFSIZE_INTS = 8192            # PacketMaster.c:44
MAX_EVENTS_PER_SEC = 2500      # PacketMaster.c:55
EOS_WORD = np.uint64(0xFFFFFFFFFFFFFFFF)


# ---------------------------------------------------------------- wire format
def words_to_wire(words):
    """Inverse of the receiver: u64 words (length multiple of 8192) -> bytes in
    PulseServer order: per bundle a 32 KiB low-order block then a 32 KiB
    high-order block, each entry big-endian (PulseServer.c:338-352)."""
    words = np.asarray(words, dtype=np.uint64)
    assert words.size % BUFSIZE_INTS == 0
    w = words.reshape(-1, BUFSIZE_INTS)
    lo = (w & np.uint64(0xFFFFFFFF)).astype('>u4')
    hi = (w >> np.uint64(32)).astype('>u4')
    out = np.empty((w.shape[0], 2, BUFSIZE_INTS), dtype='>u4')
    out[:, 0, :] = lo
    out[:, 1, :] = hi
    return out.tobytes()


def wire_to_words(buf):
    """packet = ntohl(hi[j])<<32 | ntohl(lo[j])   (PacketMaster.c:286-287,306)."""
    a = np.frombuffer(buf, dtype='>u4').reshape(-1, 2, BUFSIZE_INTS)
    lo = a[:, 0, :].astype(np.uint64)
    hi = a[:, 1, :].astype(np.uint64)
    return ((hi << np.uint64(32)) | lo).reshape(-1)


# ---------------------------------------------------------------- bitfields
def pack_word(ch, peak, p1, base, ts):
    """Inverse of unpack_fields."""
    ch, peak, p1, base, ts = (np.asarray(x, dtype=np.uint64) for x in (ch, peak, p1, base, ts))
    return ((ch & np.uint64(0xFF)) << np.uint64(56)) | ((peak & np.uint64(0xFFF)) << np.uint64(44)) | \
           ((p1 & np.uint64(0xFFF)) << np.uint64(32)) | ((base & np.uint64(0xFFF)) << np.uint64(20)) | \
           (ts & np.uint64(0xFFFFF))


def unpack_fields(words):
    """ROACH_Pulses.py:805-811: hi = word>>32, lo = word&0xffffffff;
    ch = hi/2**24 ; peak = (hi>>12)%2**12 ; p1 = hi%2**12 ;
    base = (lo>>20)%2**12 ; ts = lo%2**20."""
    w = np.asarray(words, dtype=np.uint64)
    hi = w >> np.uint64(32)
    lo = w & np.uint64(0xFFFFFFFF)
    ch = (hi // np.uint64(2 ** 24)).astype(np.uint16)
    peak = ((hi >> np.uint64(12)) % np.uint64(2 ** 12)).astype(np.uint16)
    p1 = (hi % np.uint64(2 ** 12)).astype(np.uint16)
    base = ((lo >> np.uint64(20)) % np.uint64(2 ** 12)).astype(np.uint16)
    ts = (lo % np.uint64(2 ** 20)).astype(np.uint32)
    return ch, ts, base, peak, p1


# ---------------------------------------------------------------- PacketMaster
def packetmaster_bin_literal(streams, npix_per_roach, exptime, max_events=MAX_EVENTS_PER_SEC):
    """Word-by-word restatement of PacketMaster.c:304-397 (small inputs only).

    streams: list of u64 arrays, one per roach, in arrival order.
    Returns dict(counts[exptime][R*npix] int64, lists{(sec,abs_pixel): [u64...]},
    n_eos, n_corrupt_eos, n_nonpixel, n_ignored)."""
    R = len(streams)
    counts = np.zeros((exptime, R * npix_per_roach), dtype=np.int64)
    lists = {}
    n_eos = n_corrupt = n_nonpix = n_ignored = 0
    for r, words in enumerate(streams):
        sec = 0
        plist = [0] * npix_per_roach
        photons = [[0] * max_events for _ in range(npix_per_roach)]
        for packet in np.asarray(words, dtype=np.uint64).tolist():
            adr = packet >> 56
            if sec < exptime:                                   # :327
                if adr == 255:                                  # :329
                    if packet != 0xFFFFFFFFFFFFFFFF:            # :331
                        n_corrupt += 1
                    n_eos += 1
                    for i in range(npix_per_roach):             # write_sec_data :1012-1016 (len = plist)
                        lists[(sec, r * npix_per_roach + i)] = photons[i][:plist[i]]
                        plist[i] = 0                            # :357-360
                    photons = [[0] * max_events for _ in range(npix_per_roach)]
                    sec += 1                                    # :362
                else:
                    if adr < npix_per_roach:                    # :371
                        photons[adr][plist[adr]] = packet       # :373-374
                        if plist[adr] < max_events - 1:         # :375
                            plist[adr] += 1
                            counts[sec][r * npix_per_roach + adr] += 1   # :378-379
                    else:
                        n_nonpix += 1                           # :382-386
            else:
                n_ignored += 1
    return dict(counts=counts, lists=lists, n_eos=n_eos, n_corrupt_eos=n_corrupt,
                n_nonpixel=n_nonpix, n_ignored=n_ignored)


def packetmaster_bin(streams, npix_per_roach, exptime, max_events=MAX_EVENTS_PER_SEC,
                     want_lists=False):
    """Vectorised equivalent of packetmaster_bin_literal.

    Lists (if requested) are returned flat: (sorted_words, offsets[exptime*R*npix+1])
    with key = sec*(R*npix) + abs_pixel, arrival order kept inside a key and each
    key truncated to max_events-1 entries (cap quirk PacketMaster.c:373-380)."""
    R = len(streams)
    npix = R * npix_per_roach
    raw = np.zeros(exptime * npix, dtype=np.int64)
    n_eos = n_corrupt = n_nonpix = n_ignored = 0
    keys_all, words_all = [], []
    for r, words in enumerate(streams):
        w = np.asarray(words, dtype=np.uint64)
        adr = (w >> np.uint64(56)).astype(np.int64)
        is_eos = adr == 255
        sec = np.cumsum(is_eos) - is_eos            # seconds closed BEFORE this word
        live = sec < exptime
        n_ignored += int((~live).sum())
        n_eos += int((is_eos & live).sum())
        n_corrupt += int((is_eos & live & (w != EOS_WORD)).sum())
        n_nonpix += int((live & ~is_eos & (adr >= npix_per_roach)).sum())
        ok = live & ~is_eos & (adr < npix_per_roach)
        key = sec[ok] * npix + r * npix_per_roach + adr[ok]
        raw += np.bincount(key, minlength=exptime * npix)
        if want_lists:
            keys_all.append(key)
            words_all.append(w[ok])
    counts = np.minimum(raw, max_events - 1).reshape(exptime, npix)
    out = dict(counts=counts, raw_counts=raw.reshape(exptime, npix), n_eos=n_eos,
               n_corrupt_eos=n_corrupt, n_nonpixel=n_nonpix, n_ignored=n_ignored)
    if want_lists:
        keys = np.concatenate(keys_all) if keys_all else np.zeros(0, np.int64)
        ww = np.concatenate(words_all) if words_all else np.zeros(0, np.uint64)
        order = np.argsort(keys, kind='stable')
        keys, ww = keys[order], ww[order]
        start = np.searchsorted(keys, np.arange(exptime * npix), side='left')
        rank = np.arange(keys.size) - start[keys]
        keep = rank < (max_events - 1)
        ww = ww[keep]
        offsets = np.concatenate([[0], np.cumsum(counts.reshape(-1))])
        out['list_words'] = ww
        out['list_offsets'] = offsets
    return out


def merged_list(streams, npix_per_roach, exptime, sec0=None):
    """Time-ordered merged photon list ("sorted by (sec, roach, ts)"): what the per-second flush
    of PacketMaster.c:316-342 sees before the per-pixel split.  Every valid pixel word (adr < npix_per_roach, second
    < exptime; the second of a word = end-of-second words seen before it in its roach stream, PacketMaster.c:304-342),
    key = sec * R + roach, stream order inside a key.  Returns (words, offsets[exptime*R + 1])."""
    R = len(streams)
    keys_all, words_all = [], []
    for r, words in enumerate(streams):
        w = np.asarray(words, dtype=np.uint64)
        adr = (w >> np.uint64(56)).astype(np.int64)
        is_eos = adr == 255
        sec = np.cumsum(is_eos) - is_eos + (0 if sec0 is None else int(sec0[r]))
        ok = (sec < exptime) & ~is_eos & (adr < npix_per_roach)
        keys_all.append(sec[ok] * R + r)
        words_all.append(w[ok])
    keys = np.concatenate(keys_all) if keys_all else np.zeros(0, np.int64)
    ww = np.concatenate(words_all) if words_all else np.zeros(0, np.uint64)
    order = np.argsort(keys, kind='stable')
    offsets = np.concatenate([[0], np.cumsum(np.bincount(keys, minlength=exptime * R))])
    return ww[order], offsets


def quicklook_image(counts_sec, pixel_adr):
    """write_sec_data PacketMaster.c:1029-1045: image[row][col] =
    photon_counts[sec][pixel_adr[row][col]] as uint16."""
    return np.asarray(counts_sec)[np.asarray(pixel_adr)].astype(np.uint16)


def quicklook_skysub(counts, pixel_adr, tstart, tend):
    """pulses.QuickLook (ReadoutControls/lib/pulses.py:210-236): image[i][j] += len(photons[k]) for k in
    xrange(tstart, tend) of the pixel the beammap names at [i][j] (the stored list of a second holds photon_counts
    entries, i.e. the capped count), skysub = float32(image - median(image)).  counts: capped [sec][pixel]."""
    adr = np.asarray(pixel_adr)
    image = np.zeros(adr.shape)
    for k in range(tstart, tend):
        image += np.asarray(counts[k])[adr]
    return np.float32(image - np.median(image))


# ---------------------------------------------------------------- histograms
def pixel_field_hist(streams, npix_per_roach, exptime, field='peak', bin_lut=None, n_bins=4096,
                     max_events=None):
    """Per-pixel histogram hist[R*npix][n_bins] of one 12-bit field of every live,
    valid photon word (all seconds < exptime summed).  bin_lut (4096 entries,
    values >= n_bins mean 'out of range') maps the raw field to a bin; identity
    when None.  This is the [n_pix x n_bins] product of SURVEY 2.2 K6 and the
    image_Worker 'data.bin' layout (ArconsDashboard.py:1331-1337) when n_bins=10."""
    R = len(streams)
    npix = R * npix_per_roach
    if bin_lut is None:
        bin_lut = np.arange(4096)
    bin_lut = np.asarray(bin_lut, dtype=np.int64)
    hist = np.zeros(npix * n_bins, dtype=np.int64)
    for r, words in enumerate(streams):
        w = np.asarray(words, dtype=np.uint64)
        ch, ts, base, peak, p1 = unpack_fields(w)
        adr = ch.astype(np.int64)
        is_eos = adr == 255
        sec = np.cumsum(is_eos) - is_eos
        ok = (sec < exptime) & ~is_eos & (adr < npix_per_roach)
        f = dict(peak=peak, base=base, p1=p1)[field].astype(np.int64)
        b = bin_lut[f]
        ok &= b < n_bins
        key = (r * npix_per_roach + adr[ok]) * n_bins + b[ok]
        hist += np.bincount(key, minlength=npix * n_bins)
    return hist.reshape(npix, n_bins)


def deg_bin_lut(n_bins=40, rng=(-150.0, 10.0)):
    """raw Fix12_9 field -> np.histogram bin index for the degree histograms of
    readPulses (ROACH_Pulses.py:852-860, 885-889).  Out of range -> n_bins."""
    raw = np.arange(4096, dtype='float')
    deg = (raw / 2.0 ** 9 - 4.0) * 180.0 / np.pi
    lut = np.full(4096, n_bins, dtype=np.int64)
    edges = np.histogram_bin_edges(deg, n_bins, range=rng)
    for v in range(4096):
        h, _ = np.histogram(deg[v:v + 1], n_bins, range=rng)
        nz = np.nonzero(h)[0]
        if nz.size:
            lut[v] = nz[0]
    return lut, edges


# ---------------------------------------------------------------- readPulses
def read_pulses(bram0, bram1, addr_pairs, sel_ch=0):
    """ROACH_Pulses.py:782-889 without the GUI / KATCP.

    bram0/bram1: lists (one per step) of 4*2**14-byte big-endian BRAM images
    (low / high 32 bits); addr_pairs: list of (addr0, addr1) write pointers.
    NOTE the reference only appends p1 in the wrap-around branch (:820,:828) --
    kept as is (`p1_wrap_only`)."""
    scale_to_degrees = 360. / 2 ** 12 * 4 / np.pi
    channel_count = [0] * 256
    p1 = [[] for _ in range(256)]
    timestamp = [[] for _ in range(256)]
    baseline = [[] for _ in range(256)]
    peaks = [[] for _ in range(256)]
    totals = []
    for b0, b1, (addr0, addr1) in zip(bram0, bram1, addr_pairs):
        lo = np.frombuffer(b0, dtype='>u4').astype(np.int64)
        hi = np.frombuffer(b1, dtype='>u4').astype(np.int64)
        if addr1 >= addr0:
            idx, wrap = list(range(addr0, addr1)), False
            totals.append(addr1 - addr0)
        else:
            idx, wrap = list(range(addr0, 2 ** 14)) + list(range(0, addr1)), True
            totals.append(addr1 + 2 ** 14 - addr0)
        for n in idx:
            raw_data_1 = int(hi[n])
            raw_data_0 = int(lo[n])
            ch = raw_data_1 // 2 ** 24
            channel_count[ch] += 1
            if wrap:
                p1[ch].append((raw_data_1 % 2 ** 12 - 2 ** 11) * scale_to_degrees)
            timestamp[ch].append(raw_data_0 % 2 ** 20)
            baseline[ch].append((raw_data_0 >> 20) % 2 ** 12)
            peaks[ch].append((raw_data_1 >> 12) % 2 ** 12)
    ch = sel_ch
    base = np.array(baseline[ch], dtype='float')
    base = base / 2.0 ** 9 - 4.0
    base = base * 180.0 / np.pi
    times = np.array(timestamp[ch], dtype='float') / 1e6
    peaksCh = np.array(peaks[ch], dtype='float')
    peaksCh = peaksCh / 2.0 ** 9 - 4.0
    peaksCh = peaksCh * 180.0 / np.pi
    peaksSubBase = peaksCh - base
    r = (-150, 10)
    nBin = 40
    hgBase, bins = np.histogram(base, nBin, range=r, density=False)
    hgPeak, bins = np.histogram(peaksCh, nBin, range=r, density=False)
    hgPeakSubBase, bins = np.histogram(peaksSubBase, nBin, range=r, density=False)
    return dict(channel_count=np.array(channel_count), timestamp=timestamp, baseline=baseline,
                peaks=peaks, p1=p1, total_counts=totals, base_deg=base, times=times,
                peak_deg=peaksCh, peak_sub_base=peaksSubBase, hgBase=hgBase, hgPeak=hgPeak,
                hgPeakSubBase=hgPeakSubBase, bins=bins)
