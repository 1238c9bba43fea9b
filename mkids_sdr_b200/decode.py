"""Host side of the photon-word decode / binning path (K6).

Mirrors, on the GPU, the receive/bin loop of the reference's native receiver
(DataReadout/ReadoutControls/lib/PacketMaster.c:245-405) and the photon read-out of the
channelizer GUI (DataReadout/ChannelizerControls/ROACH_Pulses.py:782-889).
"""
import ctypes

import numpy as np

from . import _lib

BUFSIZE = 32768                # PacketMaster.c:42
BUFSIZE_INTS = 8192            # PacketMaster.c:44
MAX_EVENTS_PER_SEC = 2500      # PacketMaster.c:55
FIELD_SHIFT = {'peak': 44, 'p1': 32, 'base': 20, None: -1}


def _seg_arrays(seg_offset, seg_roach, seg_sec):
    off = np.ascontiguousarray(seg_offset, dtype=np.int64)
    roach = np.ascontiguousarray(seg_roach, dtype=np.int32)
    sec = np.ascontiguousarray(seg_sec if seg_sec is not None else np.zeros(len(roach)), dtype=np.int32)
    assert off.size == roach.size + 1 == sec.size + 1
    return off, roach, sec


class PhotonDecoder:
    """Per-(second,pixel) counts + per-pixel pulse-height histogram accumulator.

    counts_raw / hist live in HBM (u32) and are accumulated across calls, so a stream can
    be fed bundle by bundle, file chunk by file chunk, or sharded across GPUs and summed
    (`counts_raw`/`hist` are plain sums; the 2500-event cap quirk of PacketMaster.c:373-380
    is applied by `counts()` after any reduction)."""

    def __init__(self, n_roaches, npix_per_roach, exptime, max_events=MAX_EVENTS_PER_SEC, hist_field='peak',
                 n_bins=4096, bin_lut=None, ctx=None, counts_buf=None, hist_buf=None):
        self.ctx = ctx or _lib.default_context()
        self.n_roaches, self.npix_per_roach, self.exptime = int(n_roaches), int(npix_per_roach), int(exptime)
        self.max_events = int(max_events)
        self.n_pix = self.n_roaches * self.npix_per_roach
        self.hist_field = hist_field
        self.n_bins = int(n_bins) if hist_field else 0
        self._lut_dev = None
        if hist_field and bin_lut is not None:
            lut = np.ascontiguousarray(np.minimum(np.asarray(bin_lut), 65535), dtype=np.uint16)
            assert lut.size == 4096
            self._lut_dev = self.ctx.to_device(lut)
        self.cfg = _lib.DecodeCfg(self.n_roaches, self.npix_per_roach, self.exptime, self.max_events,
                                  FIELD_SHIFT[hist_field], self.n_bins,
                                  self._lut_dev.ptr if self._lut_dev else None)
        # counts_buf / hist_buf: caller-owned device buffers (e.g. torch tensors to all-reduce with NCCL)
        self.counts_dev = counts_buf if counts_buf is not None else self.ctx.alloc(self.exptime * self.n_pix * 4).zero()
        if hist_field:
            self.hist_dev = hist_buf if hist_buf is not None else self.ctx.alloc(self.n_pix * self.n_bins * 4).zero()
        else:
            self.hist_dev = None
        self.stats = _lib.DecodeStats()
        self.sec = np.zeros(self.n_roaches, dtype=np.int32)     # seconds closed per roach stream

    def reset(self):
        c = self.ctx
        c._check(c.lib.mkid_memset(c.h, _lib.ptr(self.counts_dev), 0, self.exptime * self.n_pix * 4))
        if self.hist_dev is not None:
            c._check(c.lib.mkid_memset(c.h, _lib.ptr(self.hist_dev), 0, self.n_pix * self.n_bins * 4))
        self.stats = _lib.DecodeStats()
        self.sec[:] = 0

    # ------------------------------------------------------------------ feeding
    def decode_words(self, words, seg_offset, seg_roach, seg_sec=None, n_words=None, want_stats=True, want_sec=True):
        """words: u64 array (host numpy, torch tensor or DeviceBuffer).  Returns seg_sec_out (None with
        want_sec=False; with want_stats=False as well the call does not synchronise)."""
        off, roach, sec = _seg_arrays(seg_offset, seg_roach, seg_sec)
        if n_words is None:
            n_words = int(off[-1])
        sec_out = np.zeros(roach.size, dtype=np.int32) if want_sec else None
        c = self.ctx
        c._check(c.lib.mkid_decode_words(c.h, _lib.ptr(words), n_words, _lib.ptr(off), _lib.ptr(roach), _lib.ptr(sec),
                                         _lib.ptr(sec_out), roach.size, ctypes.byref(self.cfg),
                                         _lib.ptr(self.counts_dev), _lib.ptr(self.hist_dev),
                                         ctypes.addressof(self.stats) if want_stats else None))
        return sec_out

    def decode_words_seg(self, words, seg_start, seg_len, seg_roach, seg_sec=None, n_words=None, want_stats=True):
        """Explicit (start, length) segments, e.g. the per-board regions Channelizer.process fills."""
        start = np.ascontiguousarray(seg_start, dtype=np.int64)
        ln = np.ascontiguousarray(seg_len, dtype=np.int64)
        roach = np.ascontiguousarray(seg_roach, dtype=np.int32)
        sec = np.ascontiguousarray(seg_sec if seg_sec is not None else np.zeros(roach.size), dtype=np.int32)
        if n_words is None:
            n_words = int((start + ln).max()) if start.size else 0
        sec_out = np.zeros(roach.size, dtype=np.int32)
        c = self.ctx
        c._check(c.lib.mkid_decode_words_seg(c.h, _lib.ptr(words), n_words, _lib.ptr(start), _lib.ptr(ln), _lib.ptr(roach),
                                             _lib.ptr(sec), _lib.ptr(sec_out), roach.size, ctypes.byref(self.cfg),
                                             _lib.ptr(self.counts_dev), _lib.ptr(self.hist_dev),
                                             ctypes.addressof(self.stats) if want_stats else None))
        return sec_out

    def decode_words_dev(self, words, seg_start, seg_cap, seg_len_dev, seg_roach, sec_in_dev, sec_out_dev, n_words):
        """Asynchronous chaining behind a producer on the same GPU: segment lengths and the carried second
        counters live in device memory (addresses / DeviceBuffers); nothing is copied back, no synchronisation."""
        start = np.ascontiguousarray(seg_start, dtype=np.int64)
        cap = np.ascontiguousarray(seg_cap, dtype=np.int64)
        roach = np.ascontiguousarray(seg_roach, dtype=np.int32)
        c = self.ctx
        c._check(c.lib.mkid_decode_words_dev(c.h, _lib.ptr(words), int(n_words), _lib.ptr(start), _lib.ptr(cap),
                                             _lib.ptr(seg_len_dev), _lib.ptr(roach), _lib.ptr(sec_in_dev),
                                             _lib.ptr(sec_out_dev), roach.size, ctypes.byref(self.cfg),
                                             _lib.ptr(self.counts_dev), _lib.ptr(self.hist_dev)))

    def decode_lists(self, words, seg_offset, seg_roach, seg_sec=None, n_words=None):
        """Decode + the per-(second, pixel) photon lists of PacketMaster (PacketMaster.c:371-380, :1012-1016).
        Returns (list_words u64, list_offsets int64 [exptime*n_pix + 1], seg_sec_out): the words of key
        k = sec*n_pix + pixel are list_words[list_offsets[k]:list_offsets[k+1]] in arrival order, at most
        max_events-1 per key.  counts_raw / stats are accumulated as in decode_words."""
        off, roach, sec = _seg_arrays(seg_offset, seg_roach, seg_sec)
        if n_words is None:
            n_words = int(off[-1])
        sec_out = np.zeros(roach.size, dtype=np.int32)
        lw = np.empty(max(n_words, 1), dtype=np.uint64)
        lo = np.empty(self.exptime * self.n_pix + 1, dtype=np.int64)
        c = self.ctx
        c._check(c.lib.mkid_decode_lists(c.h, _lib.ptr(words), n_words, _lib.ptr(off), _lib.ptr(roach), _lib.ptr(sec),
                                         _lib.ptr(sec_out), roach.size, ctypes.byref(self.cfg), _lib.ptr(self.counts_dev),
                                         _lib.ptr(lw), lw.size, _lib.ptr(lo), ctypes.addressof(self.stats)))
        return lw[:int(lo[-1])], lo, sec_out

    def decode_merged(self, words, seg_offset, seg_roach, seg_sec=None, n_words=None):
        """Decode + the time-ordered merged photon list: every valid pixel word of the seconds
        < exptime sorted by (second, roach), stream (= timestamp) order inside.  Returns (list_words u64,
        list_offsets int64 [exptime*n_roaches + 1], seg_sec_out); the words of second s from roach r are
        list_words[list_offsets[s*n_roaches + r]:list_offsets[s*n_roaches + r + 1]]."""
        off, roach, sec = _seg_arrays(seg_offset, seg_roach, seg_sec)
        if n_words is None:
            n_words = int(off[-1])
        sec_out = np.zeros(roach.size, dtype=np.int32)
        lw = np.empty(max(n_words, 1), dtype=np.uint64)
        lo = np.empty(self.exptime * self.cfg.n_roaches + 1, dtype=np.int64)
        c = self.ctx
        c._check(c.lib.mkid_decode_merged(c.h, _lib.ptr(words), n_words, _lib.ptr(off), _lib.ptr(roach), _lib.ptr(sec),
                                          _lib.ptr(sec_out), roach.size, ctypes.byref(self.cfg), _lib.ptr(self.counts_dev),
                                          _lib.ptr(lw), lw.size, _lib.ptr(lo), ctypes.addressof(self.stats)))
        return lw[:int(lo[-1])], lo, sec_out

    def decode_wire_lists(self, wire, seg_offset, seg_roach, seg_sec=None, n_bundles=None, merged=False):
        """decode_lists (merged=False) or decode_merged (merged=True) on PulseServer bundles (bytes / u8 / u32 array,
        host or device; segments in whole bundles): PacketMaster's photon lists straight from the socket buffers."""
        if isinstance(wire, (bytes, bytearray, memoryview)):
            wire = np.frombuffer(wire, dtype=np.uint8)
        off, roach, sec = _seg_arrays(seg_offset, seg_roach, seg_sec)
        if n_bundles is None:
            n_bundles = int(off[-1])
        sec_out = np.zeros(roach.size, dtype=np.int32)
        lw = np.empty(max(n_bundles * 8192, 1), dtype=np.uint64)
        lo = np.empty(self.exptime * (self.cfg.n_roaches if merged else self.n_pix) + 1, dtype=np.int64)
        c = self.ctx
        c._check(c.lib.mkid_decode_wire_lists(c.h, _lib.ptr(wire), n_bundles, _lib.ptr(off), _lib.ptr(roach), _lib.ptr(sec),
                                              _lib.ptr(sec_out), roach.size, ctypes.byref(self.cfg), _lib.ptr(self.counts_dev),
                                              1 if merged else 0, _lib.ptr(lw), lw.size, _lib.ptr(lo),
                                              ctypes.addressof(self.stats)))
        return lw[:int(lo[-1])], lo, sec_out

    def decode_wire(self, wire, seg_offset, seg_roach, seg_sec=None, n_bundles=None, want_stats=True, want_sec=True):
        """wire: PulseServer bundles (bytes / u8 / u32 array, host or device)."""
        if isinstance(wire, (bytes, bytearray, memoryview)):
            wire = np.frombuffer(wire, dtype=np.uint8)
        off, roach, sec = _seg_arrays(seg_offset, seg_roach, seg_sec)
        if n_bundles is None:
            n_bundles = int(off[-1])
        sec_out = np.zeros(roach.size, dtype=np.int32) if want_sec else None
        c = self.ctx
        c._check(c.lib.mkid_decode_wire(c.h, _lib.ptr(wire), n_bundles, _lib.ptr(off), _lib.ptr(roach), _lib.ptr(sec),
                                        _lib.ptr(sec_out), roach.size, ctypes.byref(self.cfg),
                                        _lib.ptr(self.counts_dev), _lib.ptr(self.hist_dev),
                                        ctypes.addressof(self.stats) if want_stats else None))
        return sec_out

    def feed_bundles(self, roach, wire):
        """One or more whole bundles from one roach, in arrival order (the body of the
        PacketMaster main loop, PacketMaster.c:286-397, for `ready_roach = roach`)."""
        if isinstance(wire, (bytes, bytearray, memoryview)):
            wire = np.frombuffer(wire, dtype=np.uint8)
        nb = wire.nbytes // (2 * BUFSIZE)
        assert nb * 2 * BUFSIZE == wire.nbytes, 'whole 64 KiB bundles only'
        out = self.decode_wire(wire, [0, nb], [roach], [self.sec[roach]])
        self.sec[roach] = out[0]

    def feed_streams(self, streams):
        """streams: list (one per roach) of u64 word arrays in arrival order."""
        lens = [len(s) for s in streams]
        off = np.concatenate([[0], np.cumsum(lens)])
        words = np.ascontiguousarray(np.concatenate(streams), dtype=np.uint64) if streams else np.zeros(0, np.uint64)
        out = self.decode_words(words, off, np.arange(len(streams)), self.sec[:len(streams)].copy())
        self.sec[:len(streams)] = out

    # ------------------------------------------------------------------ results
    def _download(self, buf, count):
        out = np.empty(count, dtype=np.uint32)
        c = self.ctx
        c._check(c.lib.mkid_memcpy(c.h, _lib.ptr(out), _lib.ptr(buf), out.nbytes))
        c.sync()
        return out

    def counts_raw(self):
        return self._download(self.counts_dev, self.exptime * self.n_pix).reshape(self.exptime, self.n_pix)

    def counts(self):
        """photon_counts[sec][pixel] with the cap quirk applied (PacketMaster.c:373-380)."""
        c = self.ctx
        out = np.empty(self.exptime * self.n_pix, dtype=np.uint32)
        c._check(c.lib.mkid_counts_cap(c.h, _lib.ptr(self.counts_dev), _lib.ptr(out), out.size, self.max_events))
        return out.reshape(self.exptime, self.n_pix)

    def hist(self):
        return self._download(self.hist_dev, self.n_pix * self.n_bins).reshape(self.n_pix, self.n_bins)

    def stats_dict(self):
        s = self.stats
        return dict(n_eos=s.n_eos, n_corrupt_eos=s.n_corrupt_eos, n_nonpixel=s.n_nonpixel, n_ignored=s.n_ignored,
                    n_valid=s.n_valid)

    def quicklook_image(self, sec, pixel_adr):
        """write_sec_data quick-look (PacketMaster.c:1029-1045): capped counts of second
        `sec` gathered through the beammap `pixel_adr[rows][cols]` as uint16."""
        c = self.ctx
        capped = c.alloc(self.n_pix * 4)
        c._check(c.lib.mkid_counts_cap(c.h, ctypes.c_void_p(_lib.ptr(self.counts_dev).value + sec * self.n_pix * 4),
                                       _lib.ptr(capped), self.n_pix, self.max_events))
        adr = np.ascontiguousarray(pixel_adr, dtype=np.int32)
        img = np.empty(adr.shape, dtype=np.uint16)
        c._check(c.lib.mkid_quicklook_image(c.h, _lib.ptr(capped), _lib.ptr(adr), adr.size, _lib.ptr(img)))
        capped.free()
        return img


def parse_beammap(beam_names, npix_per_roach):
    """update_beammap_names (PacketMaster.c:880-904): the beammap holds one dataset name "/r<roach>/p<pixel>/" per image
    position; pixel_adr = roach * NPIXELS_PER_ROACH + pixel (strtok on '/', then atoi past the leading letter: leading
    digits only, 0 when there are none).  beam_names: [rows][cols] strings.  Returns int32 [rows][cols]."""
    import re

    def atoi(t):
        m = re.match(r'\s*([+-]?\d+)', t)
        return int(m.group(1)) if m else 0
    names = np.asarray(beam_names, dtype=object)
    out = np.zeros(names.shape, dtype=np.int32)
    for idx, name in np.ndenumerate(names):
        toks = [t for t in str(name).split('/') if t]              # strtok skips empty tokens
        roach = atoi(toks[0][1:]) if len(toks) > 0 else 0
        pixel = atoi(toks[1][1:]) if len(toks) > 1 else 0
        out[idx] = roach * npix_per_roach + pixel
    return out


def write_quicklook_file(obs_filepath, image, sec):
    """write_quicklook_image_v2 (PacketMaster.c:679-727): `<dir of obs>/bin/<obs root>_<sec>.txt`, one image row per line,
    every value followed by a blank ("%d "); `bin/lock.<sec>` exists while the file is being written, which is what
    the dashboard's check_files polls (ArconsDashboard.py:1217-1227).  image: uint16 [rows][cols].  Returns the path."""
    import os
    obs_dir, obs_name = os.path.split(str(obs_filepath))
    root = obs_name.split('.')[0]                                   # sscanf "%[^.].h5"
    bin_dir = os.path.join(obs_dir, 'bin')
    os.makedirs(bin_dir, exist_ok=True)
    path = os.path.join(bin_dir, '%s_%d.txt' % (root, sec))
    lock = os.path.join(bin_dir, 'lock.%d' % sec)
    open(lock, 'w').close()
    img = np.asarray(image).astype(np.uint16)
    with open(path, 'wb') as f:
        for row in img:
            f.write((''.join('%d ' % v for v in row) + '\n').encode())
    os.remove(lock)
    return path


class QuickLookWriter:
    """The per-second count images PacketMaster leaves for the dashboard (write_sec_data, PacketMaster.c:1024-1045): the
    image of second s is written once EVERY roach stream has closed it, each second once, in order.  Call `flush` with
    the decoder's per-roach second counters after feeding data."""

    def __init__(self, decoder, pixel_adr, obs_filepath):
        self.dec, self.obs_filepath = decoder, obs_filepath
        self.pixel_adr = np.ascontiguousarray(pixel_adr, dtype=np.int32)
        self.next_sec = 0
        self.written = []

    def flush(self, sec_per_roach=None):
        sec = np.asarray(self.dec.sec if sec_per_roach is None else sec_per_roach)
        ready = int(min(int(sec.min()), self.dec.exptime))
        while self.next_sec < ready:
            img = self.dec.quicklook_image(self.next_sec, self.pixel_adr)
            self.written.append(write_quicklook_file(self.obs_filepath, img, self.next_sec))
            self.next_sec += 1
        return self.written


def QuickLook(decoder, pixel_adr, tstart, tend):
    """pulses.QuickLook (ReadoutControls/lib/pulses.py:210-236) without the HDF5 file and the plot: image[i][j] = sum
    over the seconds tstart <= k < tend of the length of pixel beamimage[i][j]'s photon list of second k (the lists
    PacketMaster stores hold at most max_events - 1 photons, PacketMaster.c:373-380), then the sky is taken off as the
    median of the image.  Returns what the reference hands to imshow: float32 [rows][cols]."""
    c = decoder.ctx
    adr = np.ascontiguousarray(pixel_adr, dtype=np.int32)
    rows, cols = adr.shape
    tstart, tend = int(tstart), int(tend)
    if not (0 <= tstart and tend <= decoder.exptime):
        raise IndexError('QuickLook: seconds [%d, %d) outside the exposure of %d s' % (tstart, tend, decoder.exptime))
    image = np.zeros((rows, cols), np.float64)
    if tend - tstart == 1:
        image = decoder.quicklook_image(tstart, adr).astype(np.float64)
    elif tend > tstart:                          # (make_image's sum over [t_i, t_f) is the same sum of per-second images)
        flipped = np.empty((rows, cols), np.float64)
        c._check(c.lib.mkid_dashboard_image(c.h, _lib.ptr(decoder.counts_dev), decoder.n_pix, _lib.ptr(adr), rows, cols,
                                            tstart, tend, decoder.max_events, None, None, _lib.ptr(flipped), _lib.ptr(image)))
        c.sync()
    return np.float32(image - np.median(image))


class Dashboard:
    """Headless twin of the image part of ArconsDashboard (make_image, ArconsDashboard.py:633-723) fed straight from a
    PhotonDecoder: attributes image_time, int_time, sky_subtraction, skyrate, taking_sky, skytime, skycount,
    flat_field, flatFactors, brightpix, vmin, vmax, redpix as in the reference."""

    def __init__(self, decoder, pixel_adr):
        self.dec = decoder
        self.pixel_adr = np.ascontiguousarray(pixel_adr, dtype=np.int32)
        self.nypix, self.nxpix = self.pixel_adr.shape
        self.image_time = 0
        self.int_time = 1
        self.sky_subtraction = False
        self.taking_sky = False
        self.skytime = 0
        self.skycount = np.zeros(self.pixel_adr.shape)
        self.skyrate = np.zeros(self.pixel_adr.shape)
        self.flat_field = False
        self.flatFactors = np.ones(self.pixel_adr.shape)
        self.contrast_mode = False
        self.brightpix = 1
        self.vmin, self.vmax = 0, 0
        self.redpix = None

    def make_image(self):
        """One call per second, like the dashboard's timer: returns photon_count (the displayed frame)."""
        c = self.dec.ctx
        rows, cols = self.pixel_adr.shape
        tf = self.image_time
        ti = max(tf - int(self.int_time), 0)
        if self.taking_sky:                         # :655-656 (the flipped raw image of this second)
            self.skycount += np.flipud(self.dec.quicklook_image(tf, self.pixel_adr).astype(np.float64))
        image = np.empty((rows, cols), np.float64)
        image_counts = np.empty((rows, cols), np.float64)
        sky = np.ascontiguousarray(self.skyrate, dtype=np.float64) if self.sky_subtraction else None
        flat = np.ascontiguousarray(self.flatFactors, dtype=np.float64) if self.flat_field else None
        c._check(c.lib.mkid_dashboard_image(c.h, _lib.ptr(self.dec.counts_dev), self.dec.n_pix, _lib.ptr(self.pixel_adr), rows,
                                            cols, ti, tf, self.dec.max_events, _lib.ptr(sky), _lib.ptr(flat), _lib.ptr(image),
                                            _lib.ptr(image_counts)))
        c.sync()
        if not self.contrast_mode:                  # :695-699
            # (photon_count is a flipped VIEW of image_counts in the reference: the in-place flat-field multiply
            # reaches image_counts, so the contrast limit is taken from the flat-fielded values)
            indices = np.sort((np.flipud(image) if self.flat_field else image_counts).reshape(1, -1))
            self.vmin = 0
            self.vmax = indices[0, -1 * int(self.brightpix)]
        self.redpix = np.where(image > 2000) if tf == ti else np.where(image > 2000 * (tf - ti))     # :705-708
        self.image_counts = image_counts.reshape(1, -1)
        self.image_time += 1
        if self.taking_sky and self.image_time == self.skytime:                                       # :718-721
            self.taking_sky = False
            self.skyrate = self.skycount / self.skytime
        return image


def unpack_fields(words, ctx=None):
    """ROACH_Pulses.py:805-811 on the GPU -> (ch u8, ts u32, base u16, peak u16, p1 u16)."""
    ctx = ctx or _lib.default_context()
    w = np.ascontiguousarray(words, dtype=np.uint64)
    n = w.size
    ch = np.empty(n, np.uint8); ts = np.empty(n, np.uint32)
    base = np.empty(n, np.uint16); peak = np.empty(n, np.uint16); p1 = np.empty(n, np.uint16)
    ctx._check(ctx.lib.mkid_unpack_fields(ctx.h, _lib.ptr(w), n, _lib.ptr(ch), _lib.ptr(ts), _lib.ptr(base),
                                          _lib.ptr(peak), _lib.ptr(p1)))
    return ch, ts, base, peak, p1
