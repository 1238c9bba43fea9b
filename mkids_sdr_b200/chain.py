"""The full software readout chain of one GPU: channelize -> phase -> detect -> photon words ->
decode / per-pixel binning / histogram (K4 + K5 + K6), configured from the reference's control plane.

This is the public API a user of the reference switches to for the data path:

    chain = ReadoutChain(n_boards, n_lut, fir_int, ...); chain.set_board(b, ...)      # or ReadoutChain.synthetic(...)
    out = chain.process(iq)          # iq: int16 [n_boards][n][2], host or device

Boards (ROACH streams / feedlines) are independent; several GPUs each run their own chain over
their own boards and sum the per-pixel histograms once (see dist.py).
"""
import ctypes
import os

import numpy as np

from . import _lib
from .channelizer import Channelizer, synth_adc
from .decode import PhotonDecoder
from .pulses_form import PulsesForm
from .setup_form import SetupForm

DATA_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'data')
FS = 512e6


class ReadoutChain:
    def __init__(self, n_boards, n_lut, fir_int, mean_len=20, holdoff=1000, peak_win=32, npix_per_roach=253,
                 exptime=64, n_roaches_total=None, roach0=0, hist_field='peak', n_bins=64, bin_lut=None, ctx=None,
                 counts_buf=None, hist_buf=None, want_merged=False, pipelined=False):
        self.ctx = ctx or _lib.default_context()
        # pipelined: a second context (stream) on the same GPU takes detection, decode and the merged list of batch k
        # while the channelizer kernel of batch k + 1 runs on the first one (process_async / process_stream only)
        self.pipelined = bool(pipelined)
        self.ctx2 = _lib.Context(self.ctx.device) if self.pipelined else self.ctx
        self._k = 0
        self.n_boards, self.n_lut = n_boards, n_lut
        self.roach0 = roach0
        self.chan = Channelizer(n_boards, n_lut, mean_len, holdoff, peak_win, ctx=self.ctx)
        self.chan.set_fir(fir_int)
        n_roaches_total = n_roaches_total or n_boards
        if bin_lut is None and hist_field and n_bins < 4096:
            bin_lut = np.arange(4096) * n_bins // 4096           # coarse pulse-height spectrum
        self.dec = PhotonDecoder(n_roaches_total, npix_per_roach, exptime, hist_field=hist_field, n_bins=n_bins,
                                 bin_lut=bin_lut, ctx=self.ctx2, counts_buf=counts_buf, hist_buf=hist_buf)
        if self.pipelined:
            self.ctx._check(self.ctx.lib.mkid_chan_set_pipelined(self.ctx.h, self.chan.h, 1))
        self._words_dev = None
        self._cap = 0
        # time-ordered merged photon list of every batch, device resident: merged_words_dev /
        # merged_offsets_dev (int32 [MERGE_MAX_SEC * n_boards + 1], key = local second * n_boards + board)
        self.want_merged = bool(want_merged)
        self.merged_words_dev = self.merged_offsets_dev = None
        self.sec = np.zeros(n_boards, dtype=np.int32)
        self._sec_dev = None
        self._sec_on_host = True

    def set_board(self, b, bins, I_dds, Q_dds, zero_ch=None, centers_i=None, centers_q=None, thresholds=None):
        self.chan.set_board(b, bins, I_dds, Q_dds, zero_ch, centers_i, centers_q, thresholds)

    def reset(self):
        self.join()
        self.ctx.sync()
        self.chan.reset()
        self.dec.reset()
        self.sec[:] = 0
        self._sec_on_host = True

    def process(self, iq, n=None, words_host=None):
        """One batch: n samples per board.  Photon words stay in HBM and are decoded/binned in
        place; if words_host (u64 [n_boards][cap] array, e.g. pinned) is given they are also
        returned to the host.  Returns n_words per board."""
        if n is None:
            n = iq.shape[-2]
        if not self._sec_on_host:
            self.sync_state()
        cap = self.chan.words_capacity(n)
        if self._words_dev is None or self._cap < cap:
            self._words_dev = self.ctx.alloc(self.n_boards * cap * 8)
            self._cap = cap
        (_, n_words), _ = self.chan.process(iq, n=n, detect=True, words_out=self._words_dev, words_cap=self._cap)
        start = np.arange(self.n_boards, dtype=np.int64) * self._cap
        self.sec = self.dec.decode_words_seg(self._words_dev, start, n_words.astype(np.int64),
                                             self.roach0 + np.arange(self.n_boards), self.sec,
                                             n_words=self.n_boards * self._cap, want_stats=False)
        if words_host is not None:
            c = self.ctx
            for b in range(self.n_boards):
                if n_words[b]:
                    c._check(c.lib.mkid_memcpy(c.h, _lib.ptr(words_host[b]), self._words_dev.ptr + b * self._cap * 8,
                                               int(n_words[b]) * 8))
            c.sync()
        return n_words

    def process_async(self, iq, n=None):
        """process() without any host round trip: iq in device memory; the word counts and the carried second
        counters stay on the device, so consecutive batches queue up on the stream.  `sync_state()` brings the
        host-side view (self.sec, word counts of the last batch) up to date."""
        if n is None:
            n = iq.shape[-2]
        cap = self.chan.words_capacity(n)
        if self._words_dev is None or self._cap < cap:
            self._words_dev = self.ctx.alloc(self.n_boards * cap * 8)
            self._cap = cap
        if self._sec_dev is None:
            self._sec_dev = [self.ctx.alloc(self.n_boards * 4), self.ctx.alloc(self.n_boards * 4)]
            self._sec_cur = 0
            self._sec_on_host = True
        c = self.ctx
        if self._sec_on_host:                       # host copy is the newer one: upload once
            c._check(c.lib.mkid_memcpy(c.h, _lib.ptr(self._sec_dev[self._sec_cur]), _lib.ptr(self.sec), self.n_boards * 4))
            c.sync()
            self._sec_on_host = False
        start = np.arange(self.n_boards, dtype=np.int64) * self._cap
        caps = np.full(self.n_boards, self._cap, dtype=np.int64)
        b = self.ctx2                                   # == c unless pipelined
        if self.pipelined:
            par = self._k & 1
            if self._k >= 2:
                c.wait_event(b, 40 + par)               # the detection of batch k - 2 has released this set of phase rows
            c._check(c.lib.mkid_chan_process(c.h, self.chan.h, _lib.ptr(iq), int(n), 2, None, 0, None, None))
            self.chan.t_consumed += n // 512
            c.record(42 + par)
            b.wait_event(c, 42 + par)
            b._check(b.lib.mkid_chan_detect_pending(b.h, self.chan.h, _lib.ptr(self._words_dev), int(self._cap), None))
        else:
            self.chan.process_async(iq, self._words_dev, self._cap, n=n)
        self.dec.decode_words_dev(self._words_dev, start, caps, self.chan.n_words_dev(),
                                  self.roach0 + np.arange(self.n_boards), self._sec_dev[self._sec_cur],
                                  self._sec_dev[1 - self._sec_cur], self.n_boards * self._cap)
        if self.want_merged:
            if self.merged_words_dev is None or self.merged_words_dev.nbytes < self.n_boards * self._cap * 8:
                self.merged_words_dev = c.alloc(self.n_boards * self._cap * 8)
                self.merged_offsets_dev = c.alloc((_lib.MERGE_MAX_SEC * self.n_boards + 1) * 4)
            b._check(b.lib.mkid_merge_words_dev(b.h, _lib.ptr(self._words_dev), _lib.ptr(start), _lib.ptr(caps),
                                                self.chan.n_words_dev(), _lib.ptr(self._sec_dev[self._sec_cur]),
                                                self.n_boards, ctypes.byref(self.dec.cfg), _lib.ptr(self.merged_words_dev),
                                                self.n_boards * self._cap, _lib.ptr(self.merged_offsets_dev)))
        if self.pipelined:
            b.record(40 + (self._k & 1))
        self._k += 1
        self._sec_cur = 1 - self._sec_cur

    def process_stream(self, batches, n, words_host=None, counts_host=None, adc_format='i16'):
        """Pipelined end-to-end path for a stream of HOST batches (pinned arrays): the upload of batch k+1 (second
        stream, double buffered) overlaps the processing of batch k.  adc_format 'i16': int16 [n_boards][n][2]
        (4 bytes per complex sample); 'p12': the 12-bit packed stream of mkid_adc_unpack12, uint8 [n_boards][3 n]
        (3 bytes per sample over the host link, expanded on the GPU in front of the channelizer).  After every batch the
        photon words (whole per-board capacity, words_host u64 [n_boards][cap]), the per-(second,pixel) counts
        (counts_host) and the word counts are copied back; yields the word counts of each batch."""
        c = self.ctx
        if adc_format not in ('i16', 'p12'):
            raise ValueError("adc_format must be 'i16' or 'p12'")
        p12 = adc_format == 'p12'
        nbytes = self.n_boards * n * (3 if p12 else 4)
        if getattr(self, '_iq_dev', None) is None or self._iq_dev[0].nbytes < nbytes:
            self._iq_dev = [c.alloc(nbytes), c.alloc(nbytes)]
        if p12 and (getattr(self, '_iq_exp', None) is None or self._iq_exp.nbytes < self.n_boards * n * 4):
            self._iq_exp = c.alloc(self.n_boards * n * 4)        # one buffer: unpack and channelizer share the stream
        cap = self.chan.words_capacity(n)
        nw_host = np.zeros(self.n_boards, dtype=np.int32)
        it = iter(batches)
        nxt = next(it, None)
        k = 0
        if nxt is not None:
            c.upload_async(self._iq_dev[0], nxt, nbytes, 0)
        while nxt is not None:
            cur = k & 1
            nxt = next(it, None)
            if nxt is not None:
                c.upload_async(self._iq_dev[cur ^ 1], nxt, nbytes, cur ^ 1)      # waits for that buffer's consumer
            c.upload_wait(cur)
            if p12:
                c.adc_unpack12(self._iq_dev[cur], self.n_boards * n, self._iq_exp)
                c.upload_consumed(cur)
                self.process_async(self._iq_exp, n=n)
            else:
                self.process_async(self._iq_dev[cur], n=n)
                c.upload_consumed(cur)
            self.join()
            c._check(c.lib.mkid_memcpy(c.h, _lib.ptr(nw_host), self.chan.n_words_dev(), nw_host.nbytes))
            if words_host is not None:
                for b in range(self.n_boards):
                    c._check(c.lib.mkid_memcpy(c.h, _lib.ptr(words_host[b]), self._words_dev.ptr + b * self._cap * 8,
                                               min(cap, words_host.shape[1]) * 8))
            if counts_host is not None:
                c._check(c.lib.mkid_memcpy(c.h, _lib.ptr(counts_host), _lib.ptr(self.dec.counts_dev),
                                           self.dec.exptime * self.dec.n_pix * 4))
            c.sync()
            k += 1
            yield nw_host.copy()
        self.sync_state()

    @property
    def launches(self):
        """Kernels launched through this chain's context(s)."""
        return self.ctx.launches + (self.ctx2.launches if self.pipelined else 0)

    def join(self):
        """Pipelined mode: the first context's stream waits for everything queued on the second one (call before queuing
        work on `ctx` that reads the products, e.g. the reduce over the GPUs)."""
        if self.pipelined:
            self.ctx2.record(44)
            self.ctx.wait_event(self.ctx2, 44)

    def sync_state(self):
        """After process_async calls: returns the word counts of the last batch, refreshes self.sec."""
        self.join()
        c = self.ctx
        if self._sec_dev is not None and not self._sec_on_host:
            c._check(c.lib.mkid_memcpy(c.h, _lib.ptr(self.sec), _lib.ptr(self._sec_dev[self._sec_cur]), self.n_boards * 4))
        nw = self.chan.n_words()                      # synchronises
        self._sec_on_host = True
        return nw

    # ------------------------------------------------------------------ synthetic configuration
    @staticmethod
    def synthetic_boards(n_boards, n_lut=2 ** 19, n_active=253, seed0=42, lo_freq=5.0e9, ctx=None):
        """Per board: 256 resonator frequencies on the fs/n_lut grid (n_active of them driven), the
        DDS LUT from SetupForm.define_DDS_LUT (GPU), bins from select_bins.  Returns a list of dicts."""
        boards = []
        for b in range(n_boards):
            rng = np.random.default_rng(seed0 + b)
            k = np.sort(rng.choice(np.arange(-n_lut // 2 + 4096, n_lut // 2 - 4096), 256, replace=False))
            sf = SetupForm(N_lut_entries=n_lut, multi_tone=True, ctx=ctx)
            sf.save_npz = False
            sf.lo_freq = lo_freq
            sf.dac_freqs = [lo_freq + float(v) * FS / n_lut for v in k]
            sf.attens = np.zeros(256)
            sf.define_DDS_LUT()
            zero = np.zeros(256, np.uint8)
            zero[n_active:] = 1
            boards.append(dict(tone_bins=(k % n_lut)[:n_active].astype(np.int64), bins=np.array(sf.fft_bins),
                               I_dds=sf.I_dds.astype(np.int16), Q_dds=sf.Q_dds.astype(np.int16), zero_ch=zero,
                               residuals=np.array(sf.freq_residuals)))
        return boards

    @classmethod
    def synthetic(cls, n_boards, n_lut=2 ** 19, n_active=253, seed0=42, fir='matched_30us', threshold=None,
                  ctx=None, **kw):
        """Chain over synthetic boards.  Returns (chain, boards)."""
        ctx = ctx or _lib.default_context()
        pf = PulsesForm(N_lut_entries=n_lut, ctx=ctx)
        pf.importFIRcoeffs(os.path.join(DATA_DIR, fir + '.txt'))
        pf.dac_freqs = [0.0] * n_active
        pf.zeroChannels = [0] * 256
        pf.loadFIRcoeffs()
        boards = cls.synthetic_boards(n_boards, n_lut, n_active, seed0, ctx=ctx)
        chain = cls(n_boards, n_lut, pf.fir_int, ctx=ctx, **kw)
        for b, bd in enumerate(boards):
            thr = np.full(256, -3000 if threshold is None else threshold, np.int32)
            chain.set_board(b, bd['bins'], bd['I_dds'], bd['Q_dds'], bd['zero_ch'], None, None, thr)
        chain.fir_int = pf.fir_int
        return chain, boards

    def derive_thresholds(self, boards, n=2 ** 22, seed=7):
        """loadThresholds (ROACH_Pulses.py:259-288) on a pulse-free stretch of the synthetic stream: the raw phase
        stays in HBM and all channels are histogrammed by one kernel (mkid_thresholds_from_phase)."""
        from . import triggers
        self.join()
        self.ctx.sync()
        tb = np.stack([bd['tone_bins'] for bd in boards])
        iq = self.ctx.alloc(self.n_boards * n * 4)
        synth_adc(self.n_boards, n, tb, n_lut=self.n_lut, pulse_rate=0.0, seed=seed, out=iq, ctx=self.ctx)
        T = n // 512
        ph = self.ctx.alloc(self.n_boards * T * 256 * 2)
        self.chan.reset()
        c = self.ctx
        c._check(c.lib.mkid_chan_process(c.h, self.chan.h, _lib.ptr(iq), int(n), 0, None, 0, None, _lib.ptr(ph)))
        self.chan.reset()
        thr_all, _, _ = triggers.thresholds_from_phase(ph, self.n_boards, T, min(20480, T - 64), row0=64, ctx=self.ctx)
        iq.free(); ph.free()
        out = []
        for b in range(self.n_boards):
            thr = np.where(np.asarray(boards[b]['zero_ch']).astype(bool), -25736, thr_all[b]).astype(np.int32)
            self.chan.set_thresholds(b, thr)
            out.append(thr)
        return out
