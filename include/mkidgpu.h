/* mkidgpu.h -- C ABI of the B200-native (sm_100a) hot path of the ARCONS/MKID SDR readout.
 *
 * Plain C: pointers + sizes, no torch / C++ types.  Every entry point names the
 * reference interface it replaces (paths relative to creanero/MKIDS_SDR).
 *
 * Conventions
 *   - return 0 on success, a negative MKID_E* code on failure; the message is in
 *     mkid_last_error(ctx).  The library never calls exit() (contrast
 *     DataReadout/ReadoutControls/lib/PacketMaster.c:517-522 error()).
 *   - there is NO CPU fallback: without an sm_100 device mkid_init returns MKID_ENODEV.
 *   - data pointers may be host or device memory (detected with
 *     cudaPointerGetAttributes); host buffers are staged through the context's stream,
 *     pinned host memory (mkid_host_alloc) makes those copies asynchronous.
 *   - the caller owns every buffer; the context owns only scratch memory.
 *   - one context = one GPU + one CUDA stream; a context is not thread-safe, distinct
 *     contexts are independent (mirrors "one GUI process per roach",
 *     DataReadout/ChannelizerControls/ROACH_Setup.py:35-47).
 *   - calls are asynchronous on the context stream unless they return host scalars;
 *     mkid_sync() waits for the stream.
 */
#ifndef MKIDGPU_H
#define MKIDGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MKID_OK        0
#define MKID_ENODEV   -1   /* no sm_100 CUDA device */
#define MKID_EINVAL   -2   /* bad argument */
#define MKID_ENOMEM   -3   /* allocation failed */
#define MKID_ECUDA    -4   /* CUDA runtime error (see mkid_last_error) */
#define MKID_ENCCL    -5   /* NCCL error */

typedef struct mkid_ctx mkid_ctx;

/* ------------------------------------------------------------------ context */
int  mkid_init(int device, mkid_ctx **out);
void mkid_destroy(mkid_ctx *ctx);
const char *mkid_last_error(mkid_ctx *ctx);     /* ctx may be NULL: last init error */
int  mkid_sync(mkid_ctx *ctx);
/* the CUDA stream of the context as an opaque handle (cudaStream_t) */
void *mkid_stream(mkid_ctx *ctx);
/* Ordering against a stream the library does not own (the context stream is non-blocking: work queued by PyTorch or by
 * the caller on another stream is NOT ordered against it otherwise).  mkid_wait_stream: everything queued on
 * ext_stream (cudaStream_t, NULL = the legacy default stream) so far happens before what is queued on the context
 * afterwards; mkid_stream_wait_ctx: the reverse. */
int  mkid_wait_stream(mkid_ctx *ctx, void *ext_stream);
int  mkid_stream_wait_ctx(mkid_ctx *ctx, void *ext_stream);
/* the stream of `ctx` waits for event `slot` of another context on the same GPU (recorded with mkid_event_record) */
int  mkid_stream_wait_event(mkid_ctx *ctx, mkid_ctx *owner, int32_t slot);
/* number of kernels this context has launched since creation (bench "gpu_launches") */
int64_t mkid_launch_count(mkid_ctx *ctx);
/* device timing on the context stream: record returns an event slot id (0..63) */
int  mkid_event_record(mkid_ctx *ctx, int slot);
int  mkid_event_elapsed_ms(mkid_ctx *ctx, int slot_start, int slot_stop, float *ms);
/* pinned host / device memory helpers for callers without torch */
int  mkid_host_alloc(mkid_ctx *ctx, size_t bytes, void **out);
int  mkid_host_free(mkid_ctx *ctx, void *p);
int  mkid_dev_alloc(mkid_ctx *ctx, size_t bytes, void **out);
int  mkid_dev_free(mkid_ctx *ctx, void *p);
int  mkid_memcpy(mkid_ctx *ctx, void *dst, const void *src, size_t bytes);   /* async on the stream */
int  mkid_memset(mkid_ctx *ctx, void *dst, int value, size_t bytes);
/* double-buffered host -> device uploads on a second stream, so that the copy of batch k+1 overlaps the processing of
 * batch k: upload_async copies (pinned) host memory into a device buffer tagged with `slot` (0..3), upload_wait makes the
 * context stream wait for that copy (asynchronously), upload_consumed marks the point of the context stream after which
 * the buffer of `slot` may be overwritten by the next upload_async */
int  mkid_upload_async(mkid_ctx *ctx, void *dst_dev, const void *src_host, size_t bytes, int slot);
int  mkid_upload_wait(mkid_ctx *ctx, int slot);
int  mkid_upload_consumed(mkid_ctx *ctx, int slot);
/* writes > L2-size scratch so the next timed kernel starts with a cold L2 */
int  mkid_flush_l2(mkid_ctx *ctx);
const char *mkid_version(void);

/* ------------------------------------------------------------------ photon-word decode (K6)
 * Replaces the receive/bin loop of DataReadout/ReadoutControls/lib/PacketMaster.c:286-397
 * and the per-word unpack of DataReadout/ChannelizerControls/ROACH_Pulses.py:795-832.
 *
 * A "segment" is a contiguous run of one roach's word stream: words
 * [seg_offset[i], seg_offset[i+1]) of the input belong to roach seg_roach[i] and start
 * with seg_sec[i] seconds already closed (number of end-of-second words seen before).
 * On return seg_sec_out[i] (if not NULL) holds the seconds closed at the end of the
 * segment, so a stream can be decoded in pieces (per bundle, per file chunk, per GPU).
 */
typedef struct {
    int32_t n_roaches;          /* NROACHES                               PacketMaster.c:50   */
    int32_t npix_per_roach;     /* NPIXELS_PER_ROACH                      PacketMaster.c:52   */
    int32_t exptime;            /* seconds kept; later words are ignored  PacketMaster.c:327  */
    int32_t max_events;         /* MAX_EVENTS_PER_SEC (2500)              PacketMaster.c:55   */
    int32_t hist_field_shift;   /* 44 = peak, 32 = p1, 20 = base; < 0: no pulse-height histogram */
    int32_t n_bins;             /* bins per pixel of the pulse-height histogram */
    const uint16_t *bin_lut;    /* 4096 entries raw-field -> bin (>= n_bins: dropped); NULL = identity */
} mkid_decode_cfg;

typedef struct {
    int64_t n_eos;              /* end-of-second words seen while sec < exptime        (:329) */
    int64_t n_corrupt_eos;      /* adr==255 but word != all ones, "Corrupted EOS!"     (:331) */
    int64_t n_nonpixel;         /* adr >= npix_per_roach, "Photon from non-pixel"      (:382) */
    int64_t n_ignored;          /* words after sec == exptime                          (:327) */
    int64_t n_valid;            /* photon words binned (before the 2500 cap)                  */
} mkid_decode_stats;

/* Flat little-endian u64 words.  counts_raw [exptime][n_roaches*npix_per_roach] u32 and
 * hist [n_roaches*npix_per_roach][n_bins] u32 are ACCUMULATED into (caller zeroes them):
 * counts_raw holds uncapped per-(second,pixel) counts so that partial results from
 * several segments/GPUs can be summed; mkid_counts_cap() then applies PacketMaster's
 * cap quirk (min(count, max_events-1), PacketMaster.c:373-380).  stats is accumulated too
 * (device or host pointer; host pointer => the call synchronises). */
int mkid_decode_words(mkid_ctx *ctx, const uint64_t *words, int64_t n_words,
                      const int64_t *seg_offset, const int32_t *seg_roach,
                      const int32_t *seg_sec, int32_t *seg_sec_out, int32_t n_segments,
                      const mkid_decode_cfg *cfg, uint32_t *counts_raw, uint32_t *hist,
                      mkid_decode_stats *stats);

/* Same with explicit (start, length) segments that need not be contiguous, e.g. the per-board
 * word regions mkid_chan_process fills. */
int mkid_decode_words_seg(mkid_ctx *ctx, const uint64_t *words, int64_t n_words,
                          const int64_t *seg_start, const int64_t *seg_len, const int32_t *seg_roach,
                          const int32_t *seg_sec, int32_t *seg_sec_out, int32_t n_segments,
                          const mkid_decode_cfg *cfg, uint32_t *counts_raw, uint32_t *hist,
                          mkid_decode_stats *stats);

/* Fully asynchronous variant for a producer on the same GPU (mkid_chan_process): segment i may hold up to
 * seg_cap[i] words (host array) of which the first seg_len_dev[i] (DEVICE int32 array, written by an earlier
 * kernel on the context stream) are decoded; the carried second counters are read from seg_sec_dev and written
 * to seg_sec_out_dev (DEVICE int32 arrays, distinct).  Nothing is copied to the host; the call never
 * synchronises.  Same arithmetic as mkid_decode_words_seg (PacketMaster.c:304-397). */
int mkid_decode_words_dev(mkid_ctx *ctx, const uint64_t *words, int64_t n_words,
                          const int64_t *seg_start, const int64_t *seg_cap, const int32_t *seg_len_dev,
                          const int32_t *seg_roach, const int32_t *seg_sec_dev, int32_t *seg_sec_out_dev,
                          int32_t n_segments, const mkid_decode_cfg *cfg, uint32_t *counts_raw, uint32_t *hist);

/* The per-(second, pixel) photon lists PacketMaster keeps and writes every second (photons[r][adr][plist],
 * PacketMaster.c:371-380; write_sec_data :1012-1016): list_words holds the valid photon words sorted by
 * key = sec * n_pix + pixel, in arrival order inside a key, every key truncated to max_events - 1 entries (the cap
 * quirk); list_offsets int64 [exptime * n_pix + 1] holds the start of every key (last entry = total).  list_cap:
 * capacity of list_words in words (n_words always suffices).  counts_raw (uncapped) and stats are accumulated as in
 * mkid_decode_words.  Flat words (wire format: mkid_decode_wire_lists); a range of the input (>= 1024 words) may hold at most 4 end-of-second words.
 * Synchronises. */
int mkid_decode_lists(mkid_ctx *ctx, const uint64_t *words, int64_t n_words, const int64_t *seg_offset,
                      const int32_t *seg_roach, const int32_t *seg_sec, int32_t *seg_sec_out, int32_t n_segments,
                      const mkid_decode_cfg *cfg, uint32_t *counts_raw, uint64_t *list_words, int64_t list_cap,
                      int64_t *list_offsets, mkid_decode_stats *stats);

/* mkid_decode_lists (merged == 0) or mkid_decode_merged (merged != 0) on the wire format PacketMaster receives
 * (PulseServer bundles, PulseServer.c:318-352; segments in whole bundles as in mkid_decode_wire): the photon lists of
 * PacketMaster.c:371-380 straight from the socket buffers.  list_words are host-order u64 (`packet` of
 * PacketMaster.c:305); list_cap = 8192 * n_bundles always suffices. */
int mkid_decode_wire_lists(mkid_ctx *ctx, const uint32_t *wire, int64_t n_bundles, const int64_t *seg_offset,
                           const int32_t *seg_roach, const int32_t *seg_sec, int32_t *seg_sec_out, int32_t n_segments,
                           const mkid_decode_cfg *cfg, uint32_t *counts_raw, int32_t merged, uint64_t *list_words,
                           int64_t list_cap, int64_t *list_offsets, mkid_decode_stats *stats);

/* The time-ordered merged photon list of a readout (what PacketMaster's per-second flush
 * PacketMaster.c:316-342 hands on, without the per-pixel split): every valid pixel word of the seconds < exptime,
 * sorted by key = sec * n_roaches + roach, stream order inside a key (= timestamp order, the firmware emits the
 * words of a board in time order).  No cap.  list_offsets int64 [exptime * n_roaches + 1].  Other arguments,
 * accumulation of counts_raw / stats and the limits are those of mkid_decode_lists.  Synchronises. */
int mkid_decode_merged(mkid_ctx *ctx, const uint64_t *words, int64_t n_words, const int64_t *seg_offset,
                       const int32_t *seg_roach, const int32_t *seg_sec, int32_t *seg_sec_out, int32_t n_segments,
                       const mkid_decode_cfg *cfg, uint32_t *counts_raw, uint64_t *list_words, int64_t list_cap,
                       int64_t *list_offsets, mkid_decode_stats *stats);

/* The same merged list for ONE batch behind a producer on the same GPU (the per-board word regions of
 * mkid_chan_process), fully asynchronous like mkid_decode_words_dev: segment i (in roach order) holds seg_len_dev[i]
 * (DEVICE) of at most seg_cap[i] words starting at seg_start[i]; seg_sec_dev[i] (DEVICE) = seconds closed before it.
 * Key = local second * n_segments + i, local second = end-of-second words seen before the word inside the batch
 * (at most MKID_MERGE_MAX_SEC - 1; later words join the last key).  list_words (DEVICE, capacity list_cap) receives
 * every valid pixel word (adr < npix_per_roach, absolute second < exptime) in stream order inside its key,
 * list_offsets (DEVICE int32 [MKID_MERGE_MAX_SEC * n_segments + 1]) the start of every key.  Never synchronises. */
#define MKID_MERGE_MAX_SEC 4
int mkid_merge_words_dev(mkid_ctx *ctx, const uint64_t *words, const int64_t *seg_start, const int64_t *seg_cap,
                         const int32_t *seg_len_dev, const int32_t *seg_sec_dev, int32_t n_segments,
                         const mkid_decode_cfg *cfg, uint64_t *list_words, int64_t list_cap, int32_t *list_offsets);

/* Wire format of DataReadout/ReadoutControls/lib/PulseServer.c:318-352 as received by
 * PacketMaster.c:286-287: per bundle 8192 big-endian u32 low halves then 8192 big-endian
 * u32 high halves.  Segment offsets are in BUNDLES. */
int mkid_decode_wire(mkid_ctx *ctx, const uint32_t *wire, int64_t n_bundles,
                     const int64_t *seg_offset, const int32_t *seg_roach,
                     const int32_t *seg_sec, int32_t *seg_sec_out, int32_t n_segments,
                     const mkid_decode_cfg *cfg, uint32_t *counts_raw, uint32_t *hist,
                     mkid_decode_stats *stats);

/* counts[i] = min(counts_raw[i], max_events-1)   (PacketMaster.c:373-380), in place or not */
int mkid_counts_cap(mkid_ctx *ctx, const uint32_t *counts_raw, uint32_t *counts, int64_t n,
                    int32_t max_events);

/* Field unpack of ROACH_Pulses.py:805-811 into structure-of-arrays (any output may be NULL):
 * ch = w>>56, peak = (w>>44)&0xfff, p1 = (w>>32)&0xfff, base = (w>>20)&0xfff, ts = w&0xfffff */
int mkid_unpack_fields(mkid_ctx *ctx, const uint64_t *words, int64_t n_words, uint8_t *ch,
                       uint32_t *ts, uint16_t *base, uint16_t *peak, uint16_t *p1);

/* Utils/binTools.py:50-64 reinterpretBin: two's-complement field of nBits with binaryPoint
 * fractional bits, u64 -> f64.  (Utils/bin.py:18-29 extractBin is its scalar form.) */
int mkid_reinterpret_bin(mkid_ctx *ctx, const uint64_t *values, int64_t n, int32_t n_bits,
                         int32_t binary_point, int32_t n_bits_after_end, double *out);

/* quick-look image of PacketMaster.c:1029-1045: image[i] = (uint16) counts_sec[pixel_adr[i]] */
int mkid_quicklook_image(mkid_ctx *ctx, const uint32_t *counts_sec, const int32_t *pixel_adr,
                         int32_t n_image_pixels, uint16_t *image);

/* ------------------------------------------------------------------ channelizer + pulse detection (K4, K5)
 * Software model of the firmware channelizer that the reference only parameterises
 * (the .bof/.mdl files are absent): 512-branch 2x-oversampled polyphase filter bank + FFT-512
 * -> bin select (select_bins, ROACH_Setup.py:534-550) -> DDS mix with the LUT of
 * define_DDS_LUT (ROACH_Setup.py:506-532, layout [j][(m+154)%256][s]) -> 26-tap FIR with the
 * 12-bit taps of loadFIRcoeffs (ROACH_Pulses.py:59-111) and decimation by 2 -> centre subtract
 * (loadIQcenters, ROACH_Pulses.py:948-956) -> atan2 phase in Fix16_13 (ROACH_Pulses.py:374-378)
 * -> rolling-mean threshold trigger with hold-off (pulse_triggering_v2.py:104-174, thresholds of
 * loadThresholds ROACH_Pulses.py:259-288) -> 64-bit photon words (layout ROACH_Pulses.py:805-811)
 * with one all-ones word per second boundary (PacketMaster.c:329-333).
 * The exact arithmetic is defined by oracle/channelizer.py (parity definition of the unpinned stages).
 *
 * One mkid_chan holds n_boards independent boards (feedlines) of 256 channels each and their
 * streaming state (input history, hold-off, time).  Streams are processed in calls of n samples
 * per board (n a multiple of 2048 and >= 59392); results do not depend on how a stream is cut into calls.
 */
typedef struct mkid_chan mkid_chan;

typedef struct {
    int32_t n_boards;      /* boards (ROACH streams) processed together                         */
    int32_t n_lut;         /* DDS/DAC LUT length N = sampleRate/freqRes (ROACH_Setup.py:83-84)  */
    int32_t mean_len;      /* M: baseline = mean of the previous M phase samples (meanlength), 4..32 */
    int32_t holdoff;       /* L: dead time after a trigger in us (pulselength), >= 32           */
    int32_t peak_win;      /* W: peak search window in us, <= 60                                */
    int32_t reserved;
} mkid_chan_params;

int  mkid_chan_create(mkid_ctx *ctx, const mkid_chan_params *prm, mkid_chan **out);
void mkid_chan_destroy(mkid_ctx *ctx, mkid_chan *ch);
/* FIR taps shared by all channels: c[k] = int(tap*2047) (ROACH_Pulses.py:69,88-89), 26 ints */
int  mkid_chan_set_fir(mkid_ctx *ctx, mkid_chan *ch, const int32_t *fir_int);
/* PFB prototype window, 2048 floats (default: Hamming-windowed sinc, sum = 1) */
int  mkid_chan_set_window(mkid_ctx *ctx, mkid_chan *ch, const float *h);
/* per board: fft bins [256] (select_bins), DDS LUT I/Q int16 [n_lut] in the define_DDS_LUT
 * layout, zero_ch [256] (FIR zeroed: deleted / inactive channels, ROACH_Pulses.py:65-67,99-108),
 * centres I_c,Q_c = int(c/8) [256] (ROACH_Pulses.py:949-951), raw Fix16_13 thresholds [256]
 * (capture_threshold, ROACH_Pulses.py:286).  Host pointers. */
int  mkid_chan_set_board(mkid_ctx *ctx, mkid_chan *ch, int32_t board, const int32_t *bins,
                         const int16_t *I_dds, const int16_t *Q_dds, const uint8_t *zero_ch,
                         const int32_t *centers_i, const int32_t *centers_q, const int32_t *thresholds);
int  mkid_chan_set_thresholds(mkid_ctx *ctx, mkid_chan *ch, int32_t board, const int32_t *thresholds);
/* test hook: when set (device pointer, float [n_boards][n/512][256]) the next mkid_chan_process calls
 * also store the unquantised phase in radians of the new outputs; NULL switches it off */
int  mkid_chan_set_f32_phase_out(mkid_ctx *ctx, mkid_chan *ch, float *dev);
/* restart every stream at time 0 with empty history */
int  mkid_chan_reset(mkid_ctx *ctx, mkid_chan *ch);
/* iq: int16 [n_boards][n][2] (I,Q) host or device.  words: u64 [n_boards][words_cap];
 * n_words: int32 [n_boards] (host).  phase_out (optional): int16 [n_boards][n/512][256] Fix16_13
 * phase of the n/512 new output samples.  detect = 0 skips K5 (no words).
 * n_words == NULL (iq and words in device memory, no phase_out): the call does not synchronise; the counts stay
 * on the device (mkid_chan_n_words_dev) and words beyond words_cap are dropped. */
int  mkid_chan_process(mkid_ctx *ctx, mkid_chan *ch, const int16_t *iq, int64_t n, int32_t detect,
                       uint64_t *words, int64_t words_cap, int32_t *n_words, int16_t *phase_out);
/* K5 alone on a caller-supplied phase stream (bit-exact seam test and phase-snapshot triggering):
 * phase int16 [n_boards][rows][256]; row r is absolute time t_abs0 + r; triggers are resolved for
 * rows [mean_len, rows - peak_win - 1) ; t_next int64 [n_boards][256] in/out (host). */
/* device time (ms, CUDA events on the context stream) of the channelize kernel of the last
 * mkid_chan_process call; the call synchronises on that kernel's end event */
int  mkid_chan_last_kernel_ms(mkid_ctx *ctx, mkid_chan *ch, float *ms);
/* ... summed over the last `last_n` (<= 64) mkid_chan_process calls: one synchronisation at the end of a timed region */
int  mkid_chan_kernel_ms_sum(mkid_ctx *ctx, mkid_chan *ch, int32_t last_n, float *ms_sum);
/* device pointer to the int32 [n_boards] word counts of the last mkid_chan_process call (asynchronous chaining into
 * mkid_decode_words_dev; valid until the channelizer is destroyed) */
int  mkid_chan_n_words_dev(mkid_ctx *ctx, mkid_chan *ch, const int32_t **out);
/* Two-stage pipeline over two contexts of one GPU.  mkid_chan_set_pipelined(on): consecutive process calls alternate
 * between two sets of phase rows / candidate masks.  mkid_chan_process(..., detect = 2, ...) runs the channelizer kernel
 * (with the candidate mask) only; mkid_chan_detect_pending then runs resolve / scan / emit for THAT call, on the stream of
 * the context it is given -- which may be a second context, ordered by mkid_event_record / mkid_stream_wait_event: the
 * detection of batch k then runs under the channelizer kernel of batch k + 1 (mkids_sdr_b200/chain.py does this).
 * With the switch on, the channelizer kernel leaves 4 SMs free ((SMs - 4) / n_boards chunks per board) and
 * mkid_chan_detect_pending (i) holds its kernels back by 20 us (a one-thread kernel; MKID_TAIL_DELAY_US overrides) so that
 * the channelizer kernel of the next batch takes its SMs first -- unless the previous batch emitted more than about n / 3000
 * photon words per board, which the tail could not handle on the free SMs in the channelizer's time (decided on the device) --
 * and (ii) uses the small-tile shape of the hold-off resolver (more CTAs per SM).  Results are identical in every mode.
 * words / words_cap / n_words as in mkid_chan_process. */
int  mkid_chan_set_pipelined(mkid_ctx *ctx, mkid_chan *ch, int32_t on);
int  mkid_chan_detect_pending(mkid_ctx *ctx, mkid_chan *ch, uint64_t *words, int64_t words_cap, int32_t *n_words);
/* Asynchronous calls (n_words == NULL) cannot report a word buffer that was too small: a sticky device flag records it.
 * *flag != 0: some call since the last clear produced more words than words_cap for a board (the surplus was dropped,
 * the streaming state stayed consistent).  Synchronises. */
int  mkid_chan_overflowed(mkid_ctx *ctx, mkid_chan *ch, int32_t *flag, int32_t clear);
int  mkid_chan_detect(mkid_ctx *ctx, mkid_chan *ch, const int16_t *phase, int64_t rows, int64_t t_abs0,
                      int64_t *t_next, uint64_t *words, int64_t words_cap, int32_t *n_words);
/* 12-bit packed ADC stream.  The ADC of the readout is 12 bit (the firmware data plane the channelizer models, SURVEY.md
 * section 3: ADC 512 MS/s -> FFT-512); a complex sample can travel over the host link as 3 bytes instead of an int16
 * pair: little-endian 24-bit group I[11:0] | Q[11:0] << 12, two's complement, four samples = three 32-bit words.
 * mkid_adc_unpack12 expands n_samples (a multiple of 4) packed samples into the int16 [n][2] layout mkid_chan_process
 * reads; mkid_adc_pack12 is the inverse (values outside [-2048, 2047] are clipped and counted in *n_clipped, which
 * synchronises when non-NULL).  Device pointers, asynchronous on the context stream. */
int  mkid_adc_unpack12(mkid_ctx *ctx, const void *packed, int64_t n_samples, int16_t *iq);
int  mkid_adc_pack12(mkid_ctx *ctx, const int16_t *iq, int64_t n_samples, void *packed, int64_t *n_clipped);
/* synthetic ADC stream for tests and benchmarks (replaces the ROACH ADC): per board a comb of
 * n_tones tones at fine bins tone_bin[] (f = bin*fs/n_lut), amplitude tone_amp[], phase tone_phase[],
 * each phase-modulated by exponential pulses (rate per second, decay tau_us, depth uniform in
 * [deg_lo,deg_hi] degrees), plus white Gaussian noise, clipped to 12 bits.
 * out: int16 [n_boards][n][2] device or host. */
typedef struct {
    int32_t n_tones;
    int32_t n_lut;
    float   full_scale;     /* peak amplitude of the comb in ADC counts (<= 2047)  */
    float   noise_lsb;
    float   pulse_rate;     /* per second per tone                                  */
    float   tau_us;
    float   deg_lo, deg_hi;
    uint64_t seed;
} mkid_synth_params;
int  mkid_synth_adc(mkid_ctx *ctx, const mkid_synth_params *prm, int32_t n_boards, const int32_t *tone_bin,
                    const float *tone_amp, const float *tone_phase, int64_t n, int64_t t_abs0_us, int16_t *out);

/* ------------------------------------------------------------------ snapshot decode, software triggers, thresholds
 * (the analysis scripts and the threshold derivation around the hot path) */

/* 40-bit I/Q snapshot words of DataReadout/ChannelizerControls/pulse_triggering_IQ.py:121-147: per 16 bytes two
 * samples, I = low 16 bits of the 20-bit field in bytes 6-8 / 11-13, Q = bytes 9-10 / 14-15, big-endian,
 * two's complement (twos_comp, pulse_triggering.py:22-26).  I, Q: int16 [n_bytes/8]. */
int mkid_iq_snapshot_decode(mkid_ctx *ctx, const uint8_t *buf, int64_t n_bytes, int16_t *I, int16_t *Q);
/* pulse_triggering_IQ.py:152: deg = -360*arctan2(Q-Qc, I-Ic)/(2*pi), float64 */
int mkid_phase_deg_from_iq(mkid_ctx *ctx, const int16_t *I, const int16_t *Q, int64_t n, double Ic, double Qc, double *deg);

/* Software pulse triggers on float64 phase streams (degrees), phase: [n_streams][n]:
 *   mode 0  rolling mean, pulse_triggering_v2.py:104-174 (= pulse_triggering_IQ.py:160-200): bob = start
 *           (= 100 + meanlength); hit iff |np.mean(x[bob-M:bob]) - x[bob]| > threshold; bob += holdoff
 *           (pulselength) after a hit; the scan ends when bob + tail > n (tail = pulselength).
 *   mode 1  block mean, pulse_triggering.py:114-208 and contsnapshot ROACH_Pulses.py:614-725: means of fixed
 *           blocks of mean_len samples, hit iff |mean[bob // mean_len] - x[bob]| > threshold; wrap_negative adds
 *           360 to negative samples first (pulse_triggering.py:110-112).
 * np.mean is evaluated in NumPy's float64 operation order (sum_order 0: pairwise summation of NumPy >= 1.9;
 * 1: plain left-to-right sum of the NumPy 1.6 the reference ran on), so hit lists are bit-identical to the
 * NumPy loop.  hits: int32 [n_streams][max_hits] sample indices in order; n_hits: int32 [n_streams] (may exceed
 * max_hits: the list is then truncated). */
typedef struct {
    int32_t mode;
    int32_t mean_len;       /* meanlength (mode 0) / averagelength (mode 1)              */
    int32_t start;          /* first sample tested                                      */
    int32_t holdoff;        /* samples skipped after a hit                              */
    int32_t tail;           /* the scan stops when bob + tail > n                       */
    int32_t wrap_negative;  /* mode 1 only                                              */
    int32_t sum_order;      /* 0 = pairwise (NumPy >= 1.9), 1 = sequential (NumPy 1.6)  */
    int32_t reserved;
    double  threshold;      /* phase_threshold, degrees                                 */
} mkid_trigger_cfg;
int mkid_soft_trigger(mkid_ctx *ctx, const double *phase, int32_t n_streams, int64_t n, const mkid_trigger_cfg *cfg,
                      int32_t *hits, int32_t max_hits, int32_t *n_hits);

/* loadThresholds / loadSingleThreshold (ROACH_Pulses.py:259-288) for many channels at once, on raw Fix16_13 phase
 * in DEVICE memory (e.g. the phase_out of mkid_chan_process): sample t of channel c of board b is
 * phase[b*board_stride + t*row_stride + c].  Per channel: np.histogram(bins=100) -> n = float32(counts)/sum ->
 * tot[i] = np.sum(n[:i]) -> med = edge[argmin|tot-0.5|], p5 = edge[argmin|tot-0.05|] ->
 * threshold = int(-nsigma*|med-p5|) clamped at -25736.  thr_raw int32 [n_boards][n_ch]; med, p5 (optional) float64. */
int mkid_thresholds_from_phase(mkid_ctx *ctx, const int16_t *phase, int32_t n_boards, int64_t board_stride,
                               int64_t row_stride, int32_t n_ch, int64_t n_samples, double nsigma,
                               int32_t *thr_raw, double *med, double *p5);

/* longsnapshot noise spectrum (ROACH_Pulses.py:521-537): phase_deg float64 [n_streams][n_samples];
 * nSamplesPerFFT = n_samples / n_averages (integer division, 2^20/100 = 10485 in the reference);
 * noise_db[s][k] = mean over the n_averages segments of 20*log10(|numpy.fft.fft(segment)[k]| / norm / 1e-6),
 * float64 [n_streams][nSamplesPerFFT] (norm = 50.0 in the reference). */
int mkid_noise_spectrum(mkid_ctx *ctx, const double *phase_deg, int32_t n_streams, int64_t n_samples,
                        int32_t n_averages, double norm, double *noise_db);

/* Per-pixel spectra products of the dashboard's image_Worker (ReadoutControls/ArconsDashboard.py:1282-1504) on the
 * 10-bin per-pixel spectrum darray u32 [n_pix][10] ("data.bin", :1331-1337; K6 accumulates it with n_bins = 10):
 * medians[10] = numpy.median over the pixels of every bin (:1338-1340); optional sky subtraction x - int(median)
 * (subtract_sky); pc[p] = sum of the bins (:1453-1455); me[p] = (C0*E0 + ... + C9*E9)/pc[p], and h*c/me for
 * wavelength bins (calc_mean_energy :1356-1363), float64 in the reference's operation order.
 * bin_centres: E0..E9 (host, setup_thread :1299-1322). */
int mkid_spectra_products(mkid_ctx *ctx, const uint32_t *darray, int32_t n_pix, const double *bin_centres, double hc,
                          int32_t wavelength, int32_t sky_subtraction, double *medians, int64_t *pc, double *me);

/* Matched-filter template builder: the per-pulse arithmetic of MakeTemplate
 * (DataReadout/ReadoutControls/lib/pulses.py:239-427) batched over the pulses of one resonator.  The iqpulses table
 * is two float32 arrays I, Q [n_pulses][2000] in DEVICE memory; the accept / reject comparisons on the per-pulse
 * scalars stay with the caller (mkids_sdr_b200/template.py), exactly as in the reference loop.
 *   mkid_tpl_median      numpy.median(table[:rows, :cols]) in float32 (:273-274)
 *   mkid_tpl_prepare     per pulse: I += I1m - median(I[1:900]) IN PLACE (:283-284), P1 = arctan2(Q, I), P2 =
 *                        rad2deg(unwrap(P1)) (float32), P3 = P2 - straight-line fit over samples [0,900)+[1800,2000)
 *                        (float64, :294-295) -> P3 [n_pulses][2000]; stats [n_pulses][6]: mean(P3[:100]),
 *                        mean(P3[1900:]), std(P3[:100]), max(P3[980:1050]), max(P3), and (as two int32 in the 6th
 *                        double) the first index with P3 == that peak
 *   mkid_tpl_convpeak    argmax[j] = first arg-max of numpy.convolve(kernel600, P3[j]); p3_at[j] = P3[j][1000 + argmax - 1160]
 *   mkid_tpl_accumulate  tmpl[t] += sum, in list order, of P3[pulse[i]][(t - shift[i]) mod 2000] / norm[i]
 *                        (tP += roll(P3, shift)/max(P3), :319-320, :371-372); if noise != NULL also
 *                        noise[k] += |fft(deg2rad(roll(P3, shift)[50:850]))[k]|^2 (:376) */
int mkid_tpl_median(mkid_ctx *ctx, const float *table, int32_t rows, int32_t cols, int64_t row_stride, float *out);
int mkid_tpl_prepare(mkid_ctx *ctx, float *I, float *Q, int32_t n_pulses, float I1m, float Q1m, double *P3, double *stats);
int mkid_tpl_convpeak(mkid_ctx *ctx, const double *P3, int32_t n_pulses, const double *kernel600, int32_t *argmax, double *p3_at);
int mkid_tpl_accumulate(mkid_ctx *ctx, const double *P3, const int32_t *pulse, const int32_t *shift, const double *norm,
                        int32_t n_list, double *tmpl, double *noise);

/* Dashboard image (ReadoutControls/ArconsDashboard.py:633-723 make_image) from the per-(second,pixel) counts in
 * device memory (counts_raw [exptime][n_pix], uncapped): the per-second text images PacketMaster writes are
 * image_s[i] = (uint16) min(counts_raw[s][pixel_adr[i]], max_events-1) (PacketMaster.c:1029-1045); image_counts =
 * image_{t_f} when t_f == 0 or t_f == t_i, else sum of image_s for s in [t_i, t_f) (:673-677; the current second is
 * NOT in that sum), minus skyrate*(t_f - t_i) if skyrate != NULL (:679-680); image = flipud(image_counts) times
 * flat (if != NULL, :688-689).  skyrate is indexed like the text image, flat like the flipped frame; float64
 * [rows][cols] each. */
int mkid_dashboard_image(mkid_ctx *ctx, const uint32_t *counts_raw, int32_t n_pix, const int32_t *pixel_adr, int32_t rows,
                         int32_t cols, int32_t t_i, int32_t t_f, int32_t max_events, const double *skyrate,
                         const double *flat, double *image, double *image_counts);

/* ------------------------------------------------------------------ LUT synthesis (K1-K3)
 * mkid_comb_lut replaces AppForm.freqCombLUT (ChannelizerControls/ROACH_Setup.py:416-475; twin with GUI
 * offset/scale options ROACH_Setup_DAC.py:396-455): I[t] = sum a_n cos(2 pi f_n (t+offset)/fs + phi_n),
 * Q[t] = sum a_n sin(2 pi f_n t/fs + phi_n), scale = fudge * max(|I|,|Q|) (fudge 1.1 for echo='yes',
 * 1.0 for 'no'; scale_override > 0 replaces it: keep-old / custom scale, :456-459),
 * out = int(v*32767/scale) truncated toward zero.  Every f_n must be a multiple of
 * sample_rate/n_samples (define_DAC_LUT snaps to that grid, :498): an off-grid tone returns MKID_EINVAL (the message
 * names the set and the tone) before any work is queued, I / Q / phase / scale_out are left untouched.  random_phase != 0 draws
 * phi_n = numpy.random.seed(1000); uniform(0, 2 pi) per tone (:426-429) and returns them in phase.
 * All arrays are [batch][n_tones] / [batch][n_samples]; freq/amp/phase/scale_out are host pointers. */
int mkid_random_phases(uint32_t seed, int32_t n, double *out);
int mkid_comb_lut(mkid_ctx *ctx, const double *freq_hz, const double *amp, double *phase, int32_t n_tones,
                  double sample_rate, int32_t n_samples, int32_t offset, double fudge, int32_t random_phase,
                  double scale_override, int32_t batch, int16_t *I, int16_t *Q, double *scale_out);
/* define_DDS_LUT (ROACH_Setup.py:506-532): 256 single-tone tables of n_lut/256 samples at
 * sample_rate/512*2, each normalised to its own max, scattered to
 * [j*512 + 2*((m+ch_shift)%256) + s].  offset: the GUI sample offset applied to I only
 * (ROACH_Setup_DAC.py:397,421).  resid/phase: [batch][256] host or device; I_dds / Q_dds host, or device with
 * 4-byte alignment. */
int mkid_dds_lut(mkid_ctx *ctx, const double *resid_hz, const double *phase, double sample_rate, int32_t n_lut,
                 int32_t ch_shift, int32_t offset, int32_t batch, int16_t *I_dds, int16_t *Q_dds, double *scales_out);
/* write_LUTs (ROACH_Setup.py:560-569): per sample pair 16 bytes, big-endian int16
 * q_dds[2n+1] q_dds[2n] q_dac[2n+1] q_dac[2n] i_dds[2n+1] i_dds[2n] i_dac[2n+1] i_dac[2n]; out: 8*n bytes */
int mkid_pack_dram(mkid_ctx *ctx, const int16_t *I_dac, const int16_t *Q_dac, const int16_t *I_dds,
                   const int16_t *Q_dds, int64_t n, uint8_t *out);
/* test hook: the correctly rounded double sin/cos the exact LUT paths use */
int mkid_sincos_cr(mkid_ctx *ctx, const double *x, int64_t n, double *s, double *c);

/* ------------------------------------------------------------------ multi-GPU: the one collective of the path
 * Replaces: the single aggregator of the reference, `++photon_counts[sec][ready_roach*NPIXELS_PER_ROACH+adr]`
 * (DataReadout/ReadoutControls/lib/PacketMaster.c:371-381): one process receives all roaches.  Here every rank (one
 * process per GPU) decodes its boards / packet-file chunks into its own copy of the per-pixel products
 * (counts [sec][pixel], hist [pixel][bin], plain uint32 sums) and the copies are summed ONCE over NVLink by NCCL, in
 * place, on the context stream (no host synchronisation).  Sums are order independent: identical at any GPU count.
 * Apply mkid_counts_cap afterwards for the 2500-event quirk (PacketMaster.c:373-380).
 * NCCL is bound at run time (libnccl.so.2); without it these calls return MKID_ENCCL and nothing else is affected.
 *   mkid_nccl_unique_id : rank 0 creates the 128-byte id (ncclGetUniqueId) and hands it to the others by any means
 *   mkid_nccl_init      : every rank, collectively (ncclCommInitRank); *comm is an ncclComm_t
 *   mkid_hist_allreduce : every rank ends with the sum (the dashboards of all ranks can read it)
 *   mkid_hist_reduce    : only `root` ends with the sum (PacketMaster's single writer)                         */
int  mkid_nccl_version(mkid_ctx *ctx, int32_t *version);
int  mkid_nccl_unique_id(mkid_ctx *ctx, uint8_t id_out[128]);
int  mkid_nccl_init(mkid_ctx *ctx, const uint8_t id[128], int32_t n_ranks, int32_t rank, void **comm_out);
int  mkid_nccl_destroy(mkid_ctx *ctx, void *comm);
int  mkid_hist_allreduce(mkid_ctx *ctx, void *comm, uint32_t *products_dev, size_t n);
int  mkid_hist_reduce(mkid_ctx *ctx, void *comm, uint32_t *products_dev, size_t n, int32_t root);

#ifdef __cplusplus
}
#endif
#endif /* MKIDGPU_H */
