// K6: photon-word decode + per-(second,pixel) binning + per-pixel pulse-height histogram.
//
// Replaces the receive/bin loop of DataReadout/ReadoutControls/lib/PacketMaster.c:286-397 and
// the per-word unpack of DataReadout/ChannelizerControls/ROACH_Pulses.py:795-832.
//
// HBM-bound integer work: every word is read exactly once (8 B/word).  The only sequential dependency of
// the reference loop is "seconds closed so far" (the number of end-of-second words earlier in the same
// roach stream).  It is removed like this:
//
//   * the host cuts every segment into contiguous RANGES (whole 4096-word chunks, PacketMaster.c:42-44);
//     there are about as many ranges as resident warps (148 SMs x 2 CTAs x 16 warps);
//   * decode_stream_kernel<REL>: ONE WARP streams one range with 128-bit loads (two 2 KiB batches in flight per
//     warp), bins into its own 256-entry shared-memory row, and counts seconds LOCALLY (0,1,2,... from the start
//     of its range).  When an end-of-second word closes a local second the row is written to
//     rows[range][local second] and cleared.  No warp ever waits for another warp or CTA: no barriers, no
//     look-back, no tickets.
//   * decode_commit_kernel (one CTA per 16 ranges): sum of the end-of-second totals of the earlier ranges of the
//     segment -> absolute second at the start of the range (+ seg_sec_out); rows[range][ls] summed per
//     (roach, second) -> counts[second][pixel], statistics; rows that turn out to lie beyond exptime are
//     dropped and their histogram contribution is taken back.
//   * decode_stream_kernel<ABS> (gated on a device flag, normally exits at once): ranges with more than
//     DEC_MAX_LS seconds are finished from the recorded resume point with the now known absolute second.
#include <stdlib.h>

#include <algorithm>

#include "common.cuh"

namespace {

constexpr int DEC_WARPS = 16;                      // warps per CTA, each streams its own range
constexpr int DEC_THREADS = DEC_WARPS * 32;
constexpr int DEC_CTAS_PER_SM = 2;
constexpr int DEC_CHUNK = 4096;                    // words per chunk (wire format: half a bundle)
constexpr int DEC_BUNDLE = 8192;                   // PacketMaster.c:44 BUFSIZE_INTS
constexpr int DEC_UNIT = 256;                      // flat format: ranges are whole multiples of this many words
constexpr int DEC_MAX_LS = 4;                      // local seconds kept per range on the relative pass
constexpr int DEC_ROW = 264;                       // u32 per row: 256 channels | - | corrupt EOS | words | start
constexpr int DEC_SMEM_HIST = 4096;                // smem-privatised histogram entries per CTA

// Partitioned histogram (HIST == 3): a per-pixel histogram too large for shared memory ([253][4096] u32 = 4 MB per roach)
// costs one L2 reduction per word when it is built in place (198 G reductions per second on this GPU: 20 % of the HBM
// roof).  Instead the relative pass writes a 16-bit key (pixel inside its tile << 12 | bin) per word into the bucket of
// the word's (roach, pixel tile) - staged per warp in shared memory and stored in whole 64-byte runs - and
// part_hist_kernel, one CTA per bucket, builds the tile's histogram in shared memory from 2 bytes per word.
constexpr int PART_RING = 64;                      // staged keys per (warp, tile)
constexpr int PART_FLUSH = 32;                     // keys per store: 64 bytes
constexpr int PART_CHUNK = 128;                    // keys a warp reserves per reduction on the bucket's cursor
constexpr int PART_TILES = 32;                     // at most one tile per lane
constexpr unsigned PART_EMPTY = 0xFFFFu;           // padding key (pixel-in-tile 15 does not exist: at most 8 pixels per tile)
constexpr int PART_ROW = PART_RING * 2 + 16;         // bytes per staging ring (16-byte aligned halves); the rings of the 32 tiles fill
                                                   // at the same pace: without the skew the lanes of an append would aim at one bank
constexpr int PART_SMEM_WARP = PART_TILES * PART_ROW + PART_TILES * 4;
struct PartParams {
    uint16_t *keys;        // [n_roaches * n_tiles][cap]
    uint32_t *cursor;      // [n_roaches * n_tiles] keys reserved so far (whole chunks; beyond cap: not stored)
    uint32_t cap;          // keys per bucket, a multiple of PART_CHUNK
    int n_tiles, tp;       // tiles per roach, pixels per tile
};

struct DecRange {
    long long start;       // flat: first word (relative to words); wire: first half-bundle chunk
    int n_words;           // words in the range
    int roach, seg;
    int seg_first;         // index of the first range of this range's segment
    int seg_off;           // words between the start of the segment and the start of the range
    int pad;
};
struct DecRangeOut {
    int n_ls;              // rows written on the relative pass (<= DEC_MAX_LS)
    int resume;            // < 0: range done; else first word the absolute pass must handle
};

struct DecParams {
    const uint64_t *words;     // flat format (or nullptr)
    const uint32_t *wire;      // wire format (or nullptr)
    const DecRange *ranges;
    DecRangeOut *rout;
    int32_t *eos_tot;          // [n_ranges] end-of-second words per range
    const int32_t *seg_len_dev;// [n_seg] or nullptr: actual segment lengths in device memory (<= the host-side capacity)
    const int32_t *seg_sec;    // [n_seg] seconds closed before each segment
    int32_t *seg_sec_out;      // [n_seg] ... and after it
    int n_ranges;
    uint32_t *rows;            // [n_ranges][DEC_MAX_LS][DEC_ROW]
    int32_t *base;             // [n_ranges] absolute second at the start of each range (commit -> absolute pass)
    int *flag;                 // != 0: some range needs the absolute pass
    int n_pix, npix_per_roach, exptime, n_bins, field_shift;
    const uint16_t *bin_lut;
    uint32_t *counts;          // [exptime][n_pix]
    uint32_t *hist;            // [n_pix][n_bins] or nullptr
    unsigned long long *stats; // 5 x u64: eos, corrupt eos, non-pixel, ignored, valid
    PartParams part;           // HIST == 3
};

__device__ __forceinline__ uint32_t bswap32(uint32_t x) { return __byte_perm(x, 0, 0x0123); }
__device__ __forceinline__ unsigned warp_sum(unsigned v) { return __reduce_add_sync(0xffffffffu, v); }
__device__ __forceinline__ uint2 ld_stream_u2(const uint64_t *p) {
    uint2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ uint32_t ld_stream_u1(const uint32_t *p) {
    uint32_t r;
    asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(r) : "l"(p));
    return r;
}

// One warp, one range.  J = words per lane per 128-bit load group (flat 2, wire 4), U = groups per batch.
// One warp, one range.  J = words per lane per load group (flat: one 128-bit load = 2 words; wire: one 128-bit load
// of low halves + one of high halves = 4 words); a group is 32*J consecutive words; R groups are kept in flight.
// HIST: 0 = counts only, 1 = pulse-height histogram by global reductions, 2 = through the CTA's shared-memory copy
// NEED_LO: the fast path also needs the low 32 bits of every word (histogram of the baseline field); otherwise only
// the high halves are moved into registers (the DRAM sectors are the same, the register and L1 traffic halves)
template <bool WIRE, int HIST, bool NEED_LO, bool ABS>
struct RangeDecoder {
    static constexpr int J = WIRE ? 4 : 2;
    static constexpr int G = 32 * J;                   // words per group
    // ring depth: 16-24 registers per lane in flight (the partitioning variant needs registers of its own)
    static constexpr int R = HIST == 3 ? (WIRE ? (NEED_LO ? 2 : 3) : (NEED_LO ? 2 : 4)) : (WIRE ? (NEED_LO ? 2 : 6) : (NEED_LO ? 4 : 8));
    // position of (lane, j) inside a group: flat = lane-contiguous 32/64-bit loads, wire = one 128-bit load per lane
    __device__ __forceinline__ int idx_of(int j) const { return WIRE ? 4 * lane + j : 32 * j + lane; }

    const DecParams &p;
    uint32_t *cnt;             // this warp's shared-memory row [256]: entry adr counts words of channel adr
    uint32_t cnt_s;            //   (channels >= npix_per_roach are the "non-pixel" words), as a shared address
    uint32_t hist_s;           // CTA histogram of roach cta_roach (HIST == 2), shared address
    const uint16_t *s_lut;
    uint32_t *hist_r;          // global histogram of this range's roach
    int lane, r, roach, npix, n_bins, f_sh;
    bool f_hi, use_lut, own_roach;
    int n_words;
    const uint64_t *w_flat;
    const uint32_t *w_wire;
    long long wire_chunk0;

    int sec;                   // REL: local second, ABS: absolute second
    int row_start;             // position of the first word of the current second
    int eos_total;
    bool count_only;           // REL: more than DEC_MAX_LS seconds, only count end-of-second words from here on
    bool stop;                 // ABS: exptime reached
    int resume;
    unsigned n_bad;                                    // per lane, current second
    unsigned st_eos, st_bad, st_nonpix, st_valid;      // ABS: per lane totals
    unsigned long long st_ign;
    // HIST == 3: this warp's staging rings [PART_TILES][PART_RING] u16 and append counters [PART_TILES] (shared
    // addresses), pixels per tile and 65536 / that, rounded up; lane t keeps the state of tile t: keys flushed so far, position
    // and keys left in the chunk reserved in the bucket, bucket full (the keys go straight to the histogram then)
    uint32_t ring_s, tcnt_s, tp_n, tp_inv;
    uint32_t pt_flushed, pt_left, pt_next;             // pt_next: position of the chunk reserved ahead
    uint16_t *pt_ptr;                                  // where the next run goes
    bool pt_direct;

    __device__ __forceinline__ RangeDecoder(const DecParams &p_) : p(p_) {}

    __device__ __forceinline__ void bin(uint32_t hi, uint32_t lo) {
        const uint32_t adr = hi >> 24;                                   // (adr == 255 never gets here)
        asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(cnt_s + adr * 4) : "memory");
        if (HIST) {
            if ((int)adr >= npix) return;
            const uint32_t f = ((f_hi ? hi : lo) >> f_sh) & 0xFFFu;
            const uint32_t b = use_lut ? s_lut[f] : f;
            if ((int)b < n_bins) {
                if (HIST == 3) {
                    const uint32_t t = (adr * tp_inv) >> 16, pit = adr - t * tp_n;      // adr / tp, exact for adr < 256, tp <= 8
                    uint32_t pos;
                    asm volatile("atom.shared.add.u32 %0, [%1], 1;" : "=r"(pos) : "r"(tcnt_s + t * 4) : "memory");
                    asm volatile("st.shared.u16 [%0], %1;" ::"r"(ring_s + t * PART_ROW + (pos & (PART_RING - 1)) * 2),
                                 "h"((unsigned short)((pit << 12) | b)) : "memory");
                } else if (HIST == 2 && own_roach) asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(hist_s + (adr * n_bins + b) * 4) : "memory");
                else atomicAdd(&hist_r[adr * n_bins + b], 1u);
            }
        }
    }

    // HIST == 3.  Lane t owns tile t: when 32 keys of the tile are staged they leave as one 64-byte run (four 16-byte
    // stores) into the chunk the lane holds in the bucket of (roach, t).  The chunk after it is reserved ahead (the
    // reduction on the bucket's cursor returns while this chunk fills).  A bucket that is full - a tile that takes more
    // than four times its share of the roach's words - sends the keys to the histogram directly instead.  (A first form
    // flushed tile by tile with the whole warp, state passed by shuffles: 128 instructions per word, issue-bound.)
    __device__ __forceinline__ void part_flush_own() {
        if (pt_left == 0 && !pt_direct) {
            if (pt_next > p.part.cap - PART_CHUNK) pt_direct = true;
            else {
                const int bucket = roach * p.part.n_tiles + lane;
                pt_ptr = p.part.keys + (size_t)bucket * p.part.cap + pt_next;
                pt_left = PART_CHUNK;
                pt_next = atomicAdd(&p.part.cursor[bucket], (uint32_t)PART_CHUNK);
            }
        }
        const uint32_t src = ring_s + lane * PART_ROW + (pt_flushed & (uint32_t)PART_FLUSH) * 2;      // this half of the ring
        if (!pt_direct) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                uint4 v;
                asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(src + 16 * q) : "memory");
                reinterpret_cast<uint4 *>(pt_ptr)[q] = v;
            }
            pt_ptr += PART_FLUSH; pt_left -= PART_FLUSH;
        } else {
            const uint32_t pix0 = (uint32_t)lane * tp_n;
#pragma unroll 1
            for (int q = 0; q < PART_FLUSH / 2; ++q) {
                uint32_t w;
                asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w) : "r"(src + 4 * q) : "memory");
                const uint32_t k0 = w & 0xFFFFu, k1 = w >> 16;
                if (k0 != PART_EMPTY) atomicAdd(&hist_r[(pix0 + (k0 >> 12)) * n_bins + (k0 & 0xFFFu)], 1u);
                if (k1 != PART_EMPTY) atomicAdd(&hist_r[(pix0 + (k1 >> 12)) * n_bins + (k1 & 0xFFFu)], 1u);
            }
        }
        pt_flushed += PART_FLUSH;
    }
    // after every bin() of the warp (each lane appends at most one key per call: a ring never holds more than 31 + 32)
    __device__ __forceinline__ void part_check() {
        __syncwarp();
        uint32_t c;
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(c) : "r"(tcnt_s + lane * 4) : "memory");
        if (c - pt_flushed >= (uint32_t)PART_FLUSH) part_flush_own();
        __syncwarp();
    }
    // end of a range: the keys still staged leave padded to a whole run, the rest of the chunk and the chunk reserved ahead
    // are padded as well, so that a bucket is whole chunks of keys and padding up to its cursor
    __device__ __forceinline__ void part_drain() {
        __syncwarp();
        uint32_t c;
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(c) : "r"(tcnt_s + lane * 4) : "memory");
        const uint32_t pend = c - pt_flushed;
        if (pend > 0) {
            for (uint32_t k = pend; k < (uint32_t)PART_FLUSH; ++k)
                asm volatile("st.shared.u16 [%0], %1;" ::"r"(ring_s + lane * PART_ROW + ((pt_flushed + k) & (PART_RING - 1)) * 2),
                             "h"((unsigned short)PART_EMPTY) : "memory");
            part_flush_own();
        }
        const uint4 pad = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
        for (uint32_t k = 0; k < pt_left; k += 8) *reinterpret_cast<uint4 *>(pt_ptr + k) = pad;
        if (pt_next <= p.part.cap - PART_CHUNK) {
            uint16_t *spare = p.part.keys + (size_t)(roach * p.part.n_tiles + lane) * p.part.cap + pt_next;
            for (uint32_t k = 0; k < (uint32_t)PART_CHUNK; k += 8) *reinterpret_cast<uint4 *>(spare + k) = pad;
        }
        asm volatile("st.shared.u32 [%0], %1;" ::"r"(tcnt_s + lane * 4), "r"(0u) : "memory");
        pt_flushed = 0; pt_left = 0; pt_direct = false; pt_next = 0xFFFFFFFFu;
        __syncwarp();
    }

    // the second that ends with the word at position pe (or with the range: pe = n_words - 1, closed = false)
    __device__ __forceinline__ void flush(int pe, bool closed) {
        __syncwarp();
        const unsigned nb = warp_sum(n_bad);
        n_bad = 0;
        if (!ABS) {
            uint32_t *row = p.rows + ((size_t)r * DEC_MAX_LS + sec) * DEC_ROW;
            uint4 a = reinterpret_cast<uint4 *>(cnt)[lane * 2], b = reinterpret_cast<uint4 *>(cnt)[lane * 2 + 1];
            reinterpret_cast<uint4 *>(row)[lane * 2] = a;
            reinterpret_cast<uint4 *>(row)[lane * 2 + 1] = b;
            reinterpret_cast<uint4 *>(cnt)[lane * 2] = make_uint4(0, 0, 0, 0);
            reinterpret_cast<uint4 *>(cnt)[lane * 2 + 1] = make_uint4(0, 0, 0, 0);
            if (lane == 0) {
                row[256] = 0; row[257] = nb;
                row[258] = (unsigned)(pe + 1 - row_start);
                row[259] = (unsigned)row_start;
            }
        } else {
            uint32_t *dst = p.counts + (size_t)sec * p.n_pix + (size_t)roach * npix;
            for (int i = lane; i < 256; i += 32) {
                const uint32_t v = cnt[i];
                if (v) {
                    cnt[i] = 0;
                    if (i < npix) { atomicAdd(&dst[i], v); st_valid += v; }
                    else st_nonpix += v;
                }
            }
            if (lane == 0) { st_bad += nb; st_eos += closed ? 1u : 0u; }
        }
        __syncwarp();
    }

    // an end-of-second word at position pe closes the current second (warp-uniform)
    __device__ __forceinline__ void close_second(int pe) {
        ++eos_total;
        flush(pe, true);
        ++sec;
        row_start = pe + 1;
        if (!ABS) {
            if (sec == DEC_MAX_LS) { count_only = true; resume = pe + 1; }
        } else if (sec >= p.exptime) {
            stop = true;
            if (lane == 0) st_ign += (unsigned long long)(n_words - (pe + 1));
        }
    }

    // ordered path for one group: lane holds words idx_of(j) at positions pos_base + idx_of(j);
    // valid: bit j set if that word exists and is to be handled
    __device__ __forceinline__ void slow_group(const uint32_t (&hi)[J], const uint32_t (&lo)[J], unsigned valid, int pos_base) {
        unsigned e = 0;
#pragma unroll
        for (int j = 0; j < J; ++j)
            if (((valid >> j) & 1u) && (hi[j] >> 24) == 255u) e |= 1u << j;
        int done = 0;
        for (;;) {
            if (stop) return;
            if (count_only) {
                unsigned c = 0;
#pragma unroll
                for (int j = 0; j < J; ++j) c += ((e >> j) & 1u) && (idx_of(j) >= done);
                eos_total += (int)warp_sum(c);
                return;
            }
            int cand = 0x7fffffff;
#pragma unroll
            for (int j = J - 1; j >= 0; --j)
                if (((e >> j) & 1u) && idx_of(j) >= done) cand = idx_of(j);
            const int pe = __reduce_min_sync(0xffffffffu, cand);
#pragma unroll
            for (int j = 0; j < J; ++j) {
                const int idx = idx_of(j);
                if (((valid >> j) & 1u) && idx >= done && idx < pe) bin(hi[j], lo[j]);
                if (HIST == 3) part_check();
            }
            if (pe == 0x7fffffff) return;
#pragma unroll
            for (int j = 0; j < J; ++j)
                if (idx_of(j) == pe && (hi[j] & lo[j]) != 0xFFFFFFFFu) ++n_bad;      // "Corrupted EOS!" PacketMaster.c:331
            close_second(pos_base + pe);
            done = pe + 1;
        }
    }

    // group with per-word guards: words at positions [g0, g0 + G) restricted to [lo_bound, n_words)
    __device__ __forceinline__ void guarded_group(int g0, int lo_bound) {
        uint32_t hi[J], lo[J];
        unsigned valid = 0;
#pragma unroll
        for (int j = 0; j < J; ++j) {
            const int pos = g0 + idx_of(j);
            hi[j] = lo[j] = 0;
            if (pos >= lo_bound && pos < n_words && pos >= 0) {
                valid |= 1u << j;
                if (WIRE) {
                    const long long hc = wire_chunk0 + (pos >> 12);
                    const uint32_t *lp = w_wire + (size_t)(hc >> 1) * (2 * DEC_BUNDLE) + (size_t)(hc & 1) * DEC_CHUNK + (pos & 4095);
                    lo[j] = bswap32(ld_stream_u1(lp));
                    hi[j] = bswap32(ld_stream_u1(lp + DEC_BUNDLE));
                } else {
                    const uint2 v = ld_stream_u2(w_flat + pos);
                    lo[j] = v.x; hi[j] = v.y;
                }
            }
        }
        slow_group(hi, lo, valid, g0);
    }

    struct Group { uint32_t hi[J], lo[NEED_LO ? J : 1]; };     // wire: still big-endian

    __device__ __forceinline__ void load_group(Group &g, int pos) const {
        if (WIRE) {
            const long long hc = wire_chunk0 + (pos >> 12);
            const uint32_t *lp = w_wire + (size_t)(hc >> 1) * (2 * DEC_BUNDLE) + (size_t)(hc & 1) * DEC_CHUNK + (pos & 4095) + 4 * lane;
            const uint4 h = ld_stream_u4(reinterpret_cast<const uint4 *>(lp + DEC_BUNDLE));
            g.hi[0] = h.x; g.hi[1] = h.y; g.hi[2] = h.z; g.hi[3] = h.w;
            if (NEED_LO) {
                const uint4 l = ld_stream_u4(reinterpret_cast<const uint4 *>(lp));
                g.lo[0] = l.x; g.lo[1] = l.y; g.lo[2 % J] = l.z; g.lo[3 % J] = l.w;
            }
        } else {
#pragma unroll
            for (int j = 0; j < J; ++j) {
                const uint64_t *q = w_flat + pos + 32 * j + lane;
                if (NEED_LO) { const uint2 v = ld_stream_u2(q); g.lo[j] = v.x; g.hi[j] = v.y; }
                else g.hi[j] = ld_stream_u1(reinterpret_cast<const uint32_t *>(q) + 1);
            }
        }
    }
    __device__ __forceinline__ bool has_eos(const Group &g) const {
        bool any = false;
#pragma unroll
        for (int j = 0; j < J; ++j) any |= WIRE ? (g.hi[j] & 0xFFu) == 0xFFu : g.hi[j] >= 0xFF000000u;
        return __any_sync(0xffffffffu, any);
    }
    __device__ __forceinline__ void bin_group(const Group &g) {
#pragma unroll
        for (int j = 0; j < J; ++j) {
            const uint32_t lo = NEED_LO ? g.lo[j] : 0u;
            if (WIRE) bin(bswap32(g.hi[j]), bswap32(lo));
            else bin(g.hi[j], lo);
            if (HIST == 3) part_check();
        }
    }

    // one revolution of the ring starting at pos; CHECK: the range may end inside.  Returns false when the ring
    // has to be left: ragged tail reached, or (force) the group at pos holds an end-of-second word.
    template <bool CHECK>
    __device__ __forceinline__ bool ring_round(Group (&ring)[R], int &pos, bool &force) {
#pragma unroll
        for (int g = 0; g < R; ++g) {
            if (CHECK && pos + G > n_words) return false;
            const Group x = ring[g];
            if (!CHECK || pos + (R + 1) * G <= n_words) load_group(ring[g], pos + R * G);
            if (has_eos(x)) { force = true; return false; }
            if (!count_only) bin_group(x);
            pos += G;
        }
        return true;
    }

    __device__ void run(int r_, const DecRange &rg, int cta_roach) {
        r = r_;
        roach = rg.roach; n_words = rg.n_words;
        if (p.seg_len_dev) n_words = max(0, min(n_words, p.seg_len_dev[rg.seg] - rg.seg_off));
        own_roach = roach == cta_roach;
        hist_r = HIST ? p.hist + (size_t)roach * npix * n_bins : nullptr;
        w_flat = WIRE ? nullptr : p.words + rg.start;
        w_wire = p.wire; wire_chunk0 = rg.start;
        sec = 0; row_start = 0; eos_total = 0; count_only = false; stop = false; resume = -1;
        n_bad = 0;
        if (HIST == 3) {
            pt_flushed = 0; pt_left = 0; pt_direct = false; pt_ptr = nullptr;
            pt_next = 0xFFFFFFFFu;
            if (lane < p.part.n_tiles && n_words > 0) pt_next = atomicAdd(&p.part.cursor[roach * p.part.n_tiles + lane], (uint32_t)PART_CHUNK);
        }
        int lo_bound = 0;
        if (ABS) {
            const DecRangeOut ro = p.rout[r];
            if (ro.resume < 0) return;
            lo_bound = ro.resume; row_start = lo_bound;
            sec = p.base[r] + DEC_MAX_LS;
            if (sec >= p.exptime) {
                if (lane == 0) st_ign += (unsigned long long)(n_words - lo_bound);
                return;
            }
        }
        // The group an absolute pass resumes in, the ragged tail and every group that holds an end-of-second word
        // go through the guarded, ordered path; everything else through the ring.
        int pos = lo_bound / G * G;
        bool force = false;
#pragma unroll 1
        while (pos < n_words && !stop) {
            if (force || pos < lo_bound || pos + G > n_words) {
                guarded_group(pos, lo_bound);
                pos += G;
                force = false;
                continue;
            }
            Group ring[R];
#pragma unroll
            for (int g = 0; g < R; ++g)
                if (pos + (g + 1) * G <= n_words) load_group(ring[g], pos + g * G);
            bool go = true;
#pragma unroll 1
            while (go && pos + 2 * R * G <= n_words) go = ring_round<false>(ring, pos, force);
#pragma unroll 1
            while (go) go = ring_round<true>(ring, pos, force);
        }
        if (HIST == 3) part_drain();
        if (!ABS) {
            int n_ls = DEC_MAX_LS;
            if (!count_only) { flush(n_words - 1, false); n_ls = sec + 1; }
            if (lane == 0) {
                p.rout[r] = DecRangeOut{n_ls, resume};
                p.eos_tot[r] = eos_total;
                if (resume >= 0) atomicOr(p.flag, 1);
            }
        } else if (!stop) {
            flush(n_words - 1, false);
        }
    }
};

template <bool WIRE, int HIST, bool NEED_LO, bool ABS>
__global__ void __launch_bounds__(DEC_THREADS, DEC_CTAS_PER_SM) decode_stream_kernel(DecParams p) {
    __shared__ __align__(16) uint32_t s_cnt[DEC_WARPS][256];
    __shared__ uint32_t s_hist[HIST == 2 ? DEC_SMEM_HIST : 1];
    __shared__ uint16_t s_lut[HIST ? 4096 : 1];
    extern __shared__ __align__(16) unsigned char s_part[];      // HIST == 3: [DEC_WARPS][PART_SMEM_WARP]
    if (ABS && *p.flag == 0) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool use_lut = HIST != 0 && p.bin_lut != nullptr;
    if (use_lut)
        for (int i = tid; i < 4096; i += DEC_THREADS) s_lut[i] = p.bin_lut[i];
    for (int i = tid; i < DEC_WARPS * 256; i += DEC_THREADS) (&s_cnt[0][0])[i] = 0;
    if (HIST == 2)
        for (int i = tid; i < DEC_SMEM_HIST; i += DEC_THREADS) s_hist[i] = 0;
    if (HIST == 3)
        for (int i = tid; i < DEC_WARPS * PART_SMEM_WARP / 4; i += DEC_THREADS) reinterpret_cast<uint32_t *>(s_part)[i] = 0u;
    __syncthreads();

    int r = blockIdx.x * DEC_WARPS + warp;
    const int cta_roach = blockIdx.x * DEC_WARPS < p.n_ranges ? p.ranges[blockIdx.x * DEC_WARPS].roach : -1;
    RangeDecoder<WIRE, HIST, NEED_LO, ABS> d(p);
    d.cnt = s_cnt[warp]; d.cnt_s = mk_smem_u32(s_cnt[warp]); d.hist_s = mk_smem_u32(s_hist); d.s_lut = s_lut; d.lane = lane;
    d.npix = p.npix_per_roach; d.n_bins = p.n_bins;
    if (HIST == 3) {
        d.ring_s = mk_smem_u32(s_part + warp * PART_SMEM_WARP);
        d.tcnt_s = d.ring_s + PART_TILES * PART_ROW;
        d.tp_n = (uint32_t)p.part.tp; d.tp_inv = (65536u + d.tp_n - 1) / d.tp_n;
    }
    d.use_lut = use_lut;
    d.f_hi = p.field_shift >= 32; d.f_sh = p.field_shift & 31;
    d.st_eos = d.st_bad = d.st_nonpix = d.st_valid = 0; d.st_ign = 0;
    for (; r < p.n_ranges; r += gridDim.x * DEC_WARPS) d.run(r, p.ranges[r], cta_roach);
    if (ABS) {
        const unsigned e = warp_sum(d.st_eos), b = warp_sum(d.st_bad), n = warp_sum(d.st_nonpix), v = warp_sum(d.st_valid);
        if (lane == 0) {
            if (e) atomicAdd(&p.stats[0], (unsigned long long)e);
            if (b) atomicAdd(&p.stats[1], (unsigned long long)b);
            if (n) atomicAdd(&p.stats[2], (unsigned long long)n);
            if (d.st_ign) atomicAdd(&p.stats[3], d.st_ign);
            if (v) atomicAdd(&p.stats[4], (unsigned long long)v);
        }
    }
    if (HIST == 2) {
        __syncthreads();
        if (cta_roach >= 0) {
            uint32_t *hist_r = p.hist + (size_t)cta_roach * p.npix_per_roach * p.n_bins;
            const int n = p.npix_per_roach * p.n_bins;
            for (int i = tid; i < n; i += DEC_THREADS) {
                const uint32_t v = s_hist[i];
                if (v) atomicAdd(&hist_r[i], v);
            }
        }
    }
}

// One CTA per DEC_COMMIT_RANGES consecutive ranges, thread t = channel t.  Absolute second at the start of every
// range (sum of the end-of-second totals of the earlier ranges of its segment), then
// rows[range][local second] -> counts[second][pixel] + statistics.  Consecutive ranges mostly lie in the same
// (roach, second): their rows are summed in a register and reach counts[] as ONE reduction per pixel (the rows
// of all ranges aimed at the same few thousand addresses would serialise in the L2 atomic units).  Rows that turn
// out to lie beyond exptime are dropped and their histogram contribution (made before the second was known)
// is taken back.
constexpr int DEC_COMMIT_RANGES = 16;
template <bool WIRE>
__global__ void __launch_bounds__(256) decode_commit_kernel(DecParams p) {
    constexpr int CR = DEC_COMMIT_RANGES;
    __shared__ unsigned long long s_st[5];
    __shared__ int s_part[8];
    __shared__ int s_base[CR], s_nls[CR], s_eos[CR], s_roach[CR], s_seg[CR], s_first[CR];
    __shared__ unsigned s_bad0[CR];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid < 5) s_st[tid] = 0;
    const int r0 = blockIdx.x * CR;
    const int n_here = min(CR, p.n_ranges - r0);
    const int f0 = p.ranges[r0].seg_first;
    int part = 0;
    for (int i = f0 + tid; i < r0; i += 256) part += p.eos_tot[i];
    part = (int)warp_sum((unsigned)part);
    if (lane == 0) s_part[warp] = part;
    if (tid < n_here) {
        const DecRange rg = p.ranges[r0 + tid];
        s_nls[tid] = p.rout[r0 + tid].n_ls;
        s_eos[tid] = p.eos_tot[r0 + tid];
        s_roach[tid] = rg.roach; s_seg[tid] = rg.seg; s_first[tid] = rg.seg_first;
        s_bad0[tid] = p.rows[((size_t)(r0 + tid) * DEC_MAX_LS) * DEC_ROW + 257];
    }
    __syncthreads();
    if (tid == 0) {
        int run = p.seg_sec[s_seg[0]];
        for (int w = 0; w < 8; ++w) run += s_part[w];
        for (int i = 0; i < n_here; ++i) {
            if (s_first[i] == r0 + i) run = p.seg_sec[s_seg[i]];          // a segment starts here
            s_base[i] = run;
            p.base[r0 + i] = run;
            run += s_eos[i];
            const bool last_of_seg = r0 + i + 1 == p.n_ranges || (i + 1 < n_here ? s_seg[i + 1] != s_seg[i] : p.ranges[r0 + i + 1].seg != s_seg[i]);
            if (last_of_seg && p.seg_sec_out) p.seg_sec_out[s_seg[i]] = run;
        }
    }
    // first rows of all ranges: independent loads, one latency
    uint32_t v0[CR];
#pragma unroll
    for (int i = 0; i < CR; ++i) v0[i] = i < n_here ? p.rows[((size_t)(r0 + i) * DEC_MAX_LS) * DEC_ROW + tid] : 0u;
    __syncthreads();

    unsigned long long eos = 0, bad = 0, ign = 0, valid = 0, nonpix = 0;
    const bool use_lut = p.bin_lut != nullptr;
    const bool f_hi = p.field_shift >= 32;
    const int f_sh = p.field_shift & 31;
    const bool is_pix = tid < p.npix_per_roach;
    uint32_t acc = 0;
    int key_sec = -1, key_roach = -1;
    auto flush_acc = [&]() {
        if (acc) atomicAdd(&p.counts[(size_t)key_sec * p.n_pix + (size_t)key_roach * p.npix_per_roach + tid], acc);
        acc = 0;
    };
#pragma unroll
    for (int i = 0; i < CR; ++i) {
        if (i >= n_here) break;
        const int base = s_base[i], nls = s_nls[i], roach = s_roach[i];
        for (int ls = 0; ls < nls; ++ls) {
            const int sec = base + ls;
            const uint32_t *row = p.rows + ((size_t)(r0 + i) * DEC_MAX_LS + ls) * DEC_ROW;
            if (sec < p.exptime) {
                const uint32_t v = ls == 0 ? v0[i] : row[tid];
                if (is_pix) {
                    if (sec != key_sec || roach != key_roach) { flush_acc(); key_sec = sec; key_roach = roach; }
                    acc += v; valid += v;
                } else if (tid < 255) nonpix += v;                // (channel 255 is never counted in a row)
                if (tid == 0) { bad += ls == 0 ? s_bad0[i] : row[257]; eos += ls < s_eos[i] ? 1 : 0; }
                continue;
            }
            const unsigned n_w = row[258], w0 = row[259];
            if (tid == 0) ign += n_w;
            if (!p.hist) continue;
            const long long start = p.ranges[r0 + i].start;
            for (unsigned w = w0 + tid; w < w0 + n_w; w += 256) {
                uint32_t hi, lo;
                if (WIRE) {
                    const long long hc = start + (w >> 12);
                    const uint32_t *lp = p.wire + (size_t)(hc >> 1) * (2 * DEC_BUNDLE) + (size_t)(hc & 1) * DEC_CHUNK + (w & 4095);
                    lo = bswap32(lp[0]); hi = bswap32(lp[DEC_BUNDLE]);
                } else {
                    const uint64_t x = p.words[start + w];
                    hi = (uint32_t)(x >> 32); lo = (uint32_t)x;
                }
                const uint32_t adr = hi >> 24;
                if ((int)adr >= p.npix_per_roach) continue;
                const uint32_t f = ((f_hi ? hi : lo) >> f_sh) & 0xFFFu;
                const uint32_t b = use_lut ? p.bin_lut[f] : f;
                if ((int)b < p.n_bins)
                    atomicAdd(&p.hist[((size_t)roach * p.npix_per_roach + adr) * p.n_bins + b], 0xFFFFFFFFu);
            }
        }
    }
    if (is_pix) flush_acc();
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        valid += __shfl_xor_sync(0xffffffffu, valid, d);
        nonpix += __shfl_xor_sync(0xffffffffu, nonpix, d);
    }
    if (lane == 0) {
        if (valid) atomicAdd(&s_st[4], valid);
        if (nonpix) atomicAdd(&s_st[2], nonpix);
    }
    if (tid == 0) {
        if (eos) atomicAdd(&s_st[0], eos);
        if (bad) atomicAdd(&s_st[1], bad);
        if (ign) atomicAdd(&s_st[3], ign);
    }
    __syncthreads();
    if (tid < 5 && s_st[tid]) atomicAdd(&p.stats[tid], s_st[tid]);
}

// HIST == 3, second step: one CTA per (roach, pixel tile) reads the tile's keys (2 bytes per word, whole chunks up to the
// cursor, padding skipped), builds the tile's [pixels][bins] histogram in shared memory and adds it to the global one: the
// tile is contiguous there and no other CTA touches it, so plain coalesced read-modify-writes do.  The cursor is cleared for
// the next call.
__global__ void __launch_bounds__(1024, 1) part_hist_kernel(PartParams q, uint32_t *hist, int npix_per_roach, int n_bins) {
    extern __shared__ __align__(16) uint32_t s_h[];
    const int bucket = blockIdx.x, roach = bucket / q.n_tiles, tile = bucket % q.n_tiles, tid = threadIdx.x;
    const int pix0 = tile * q.tp, np = min(q.tp, npix_per_roach - pix0);
    const uint32_t n = min(q.cursor[bucket], q.cap);
    const int cells = np * n_bins;
    for (int i = tid; i < cells; i += 1024) s_h[i] = 0u;
    __syncthreads();
    if (tid == 0) q.cursor[bucket] = 0u;
    if (n == 0 || np <= 0) return;
    const uint4 *src = reinterpret_cast<const uint4 *>(q.keys + (size_t)bucket * q.cap);
    const uint32_t hs = mk_smem_u32(s_h), n16 = n / 8;
    auto add2 = [&](uint32_t v) {
        const uint32_t k0 = v & 0xFFFFu, k1 = v >> 16;
        if (k0 != PART_EMPTY) asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(hs + ((k0 >> 12) * n_bins + (k0 & 0xFFFu)) * 4) : "memory");
        if (k1 != PART_EMPTY) asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(hs + ((k1 >> 12) * n_bins + (k1 & 0xFFFu)) * 4) : "memory");
    };
    uint32_t i = tid;
    for (; i + 3 * 1024 < n16; i += 4 * 1024) {          // four 16-byte loads in flight per thread
        const uint4 a = ld_stream_u4(src + i), b = ld_stream_u4(src + i + 1024), c = ld_stream_u4(src + i + 2048), d = ld_stream_u4(src + i + 3072);
        add2(a.x); add2(a.y); add2(a.z); add2(a.w);
        add2(b.x); add2(b.y); add2(b.z); add2(b.w);
        add2(c.x); add2(c.y); add2(c.z); add2(c.w);
        add2(d.x); add2(d.y); add2(d.z); add2(d.w);
    }
    for (; i < n16; i += 1024) {
        const uint4 a = ld_stream_u4(src + i);
        add2(a.x); add2(a.y); add2(a.z); add2(a.w);
    }
    __syncthreads();
    uint32_t *g = hist + ((size_t)roach * npix_per_roach + pix0) * n_bins;
    for (int j = tid; j < cells; j += 1024) {
        const uint32_t v = s_h[j];
        if (v) g[j] += v;
    }
}

// ------------------------------------------------------------------------------------------------
// Short segments chained behind a producer on the same GPU (mkid_decode_words_dev: the per-board photon words of one
// channelizer batch, ~10^4 words each): one CTA per segment does the whole job in ONE launch - block scan of the
// end-of-second words for the second of every word, direct reductions into counts / hist, statistics, carried
// second - instead of the three launches of the range machinery (45 us for 75 k words).
struct DecSeg { long long start; int cap, roach; };
constexpr int DEC_SMALL_MAX_WORDS = 1 << 18;           // total capacity up to which the dev entry point takes this path
__global__ void __launch_bounds__(1024) decode_small_kernel(DecParams p, const DecSeg *segs) {
    __shared__ int s_w[33];                              // exclusive end-of-second counts of the warps | chunk total
    __shared__ unsigned long long s_st[5];
    const int seg = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const DecSeg sg = segs[seg];
    const int len = max(0, min(sg.cap, p.seg_len_dev[seg]));
    const uint64_t *w = p.words + sg.start;
    const bool use_lut = p.bin_lut != nullptr, f_hi = p.field_shift >= 32;
    const int f_sh = p.field_shift & 31, npix = p.npix_per_roach;
    const unsigned lt = (1u << lane) - 1u;
    if (tid < 5) s_st[tid] = 0;
    int sec = p.seg_sec[seg];                            // second at the start of the chunk
    unsigned n_eos = 0, n_bad = 0, n_nonpix = 0, n_ign = 0, n_valid = 0;
    for (int base = 0; base < len; base += 1024) {
        const int i = base + tid;
        const bool valid = i < len;
        const uint64_t x = valid ? w[i] : 0ull;
        const uint32_t hi = (uint32_t)(x >> 32), lo = (uint32_t)x, adr = hi >> 24;
        const bool is_eos = valid && adr == 255u;
        const unsigned b = __ballot_sync(0xffffffffu, is_eos);
        if (lane == 0) s_w[warp] = __popc(b);
        __syncthreads();
        if (warp == 0) {
            const int v = s_w[lane];
            int incl = v;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const int a = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += a; }
            s_w[lane] = incl - v;
            if (lane == 31) s_w[32] = incl;
        }
        __syncthreads();
        const int my_sec = sec + s_w[warp] + __popc(b & lt);           // end-of-second words before this one
        sec += s_w[32];
        if (valid) {
            if (my_sec >= p.exptime) ++n_ign;
            else if (is_eos) { ++n_eos; if (x != ~0ull) ++n_bad; }
            else if ((int)adr >= npix) ++n_nonpix;
            else {
                ++n_valid;
                const size_t pix = (size_t)sg.roach * npix + adr;
                atomicAdd(&p.counts[(size_t)my_sec * p.n_pix + pix], 1u);
                if (p.hist) {
                    const uint32_t f = ((f_hi ? hi : lo) >> f_sh) & 0xFFFu;
                    const uint32_t bin = use_lut ? p.bin_lut[f] : f;
                    if ((int)bin < p.n_bins) atomicAdd(&p.hist[pix * p.n_bins + bin], 1u);
                }
            }
        }
        __syncthreads();                                 // s_w is rewritten by the next chunk
    }
    if (tid == 0 && p.seg_sec_out) p.seg_sec_out[seg] = sec;
    n_eos = warp_sum(n_eos); n_bad = warp_sum(n_bad); n_nonpix = warp_sum(n_nonpix);
    n_ign = warp_sum(n_ign); n_valid = warp_sum(n_valid);
    if (lane == 0) {
        if (n_eos) atomicAdd(&s_st[0], (unsigned long long)n_eos);
        if (n_bad) atomicAdd(&s_st[1], (unsigned long long)n_bad);
        if (n_nonpix) atomicAdd(&s_st[2], (unsigned long long)n_nonpix);
        if (n_ign) atomicAdd(&s_st[3], (unsigned long long)n_ign);
        if (n_valid) atomicAdd(&s_st[4], (unsigned long long)n_valid);
    }
    __syncthreads();
    if (tid < 5 && s_st[tid]) atomicAdd(&p.stats[tid], s_st[tid]);
}

// ------------------------------------------------------------------------------------------------
// Time-ordered merged photon list of ONE channelizer batch, device-chained like decode_small_kernel (SURVEY 8d config 4:
// "merged photon list sorted by (sec, roach, ts)"; what PacketMaster's per-second flush hands on, PacketMaster.c:316-342,
// without the per-pixel split).  Key = local second * n_segments + segment; the words of a board are already in time
// order, so the list is a stable compaction: every valid pixel word (adr < npix, absolute second < exptime) in stream
// order inside its key.  Two launches of one CTA per segment: per-key counts, then offsets + copy.
constexpr int MERGE_MAX_SEC = 4;                       // end-of-second words a batch segment may hold (later words join the last key)
struct MergeParams {
    const uint64_t *words;
    const int32_t *seg_len_dev, *seg_sec_dev;
    int n_seg, npix_per_roach, exptime;
    int32_t *cnt;                                      // [MERGE_MAX_SEC][n_seg]
    uint64_t *out;
    long long out_cap;
    int32_t *offsets;                                  // [MERGE_MAX_SEC * n_seg + 1]
};
template <bool COPY>
__global__ void __launch_bounds__(1024) merge_small_kernel(MergeParams p, const DecSeg *segs) {
    __shared__ int s_eos[33], s_val[33];
    __shared__ int s_base[MERGE_MAX_SEC], s_cnt[MERGE_MAX_SEC];
    const int seg = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const DecSeg sg = segs[seg];
    const int len = max(0, min(sg.cap, p.seg_len_dev[seg]));
    const uint64_t *w = p.words + sg.start;
    const int sec0 = p.seg_sec_dev[seg];
    const unsigned lt = (1u << lane) - 1u;
    if (tid < MERGE_MAX_SEC) s_cnt[tid] = 0;
    if (COPY && tid < MERGE_MAX_SEC) {                 // start of key (tid, seg): counts of all earlier keys
        int b = 0;
        for (int k = 0; k < tid * p.n_seg + seg; ++k) b += p.cnt[k];
        s_base[tid] = b;
    }
    if (COPY && seg == 0) {                            // the offsets table (exclusive scan of the counts)
        const int n_keys = MERGE_MAX_SEC * p.n_seg;
        if (tid == 0) { int run = 0; for (int k = 0; k < n_keys; ++k) { p.offsets[k] = run; run += p.cnt[k]; } p.offsets[n_keys] = run; }
    }
    __syncthreads();
    int ls_run = 0, val_run = 0;                       // end-of-second words / valid words before this chunk
    int done_before[MERGE_MAX_SEC];                    // valid words of this segment in keys before ls (COPY)
#pragma unroll
    for (int l = 0; l < MERGE_MAX_SEC; ++l) {
        int a = 0;
        if (COPY) for (int q = 0; q < l; ++q) a += p.cnt[q * p.n_seg + seg];
        done_before[l] = a;
    }
    for (int base = 0; base < len; base += 1024) {
        const int i = base + tid;
        const bool in = i < len;
        const uint64_t x = in ? w[i] : 0ull;
        const uint32_t adr = (uint32_t)(x >> 56);
        const bool is_eos = in && adr == 255u;
        const unsigned be = __ballot_sync(0xffffffffu, is_eos);
        if (lane == 0) s_eos[warp] = __popc(be);
        __syncthreads();
        if (warp == 0) {
            const int v = s_eos[lane];
            int incl = v;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const int a = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += a; }
            s_eos[lane] = incl - v;
            if (lane == 31) s_eos[32] = incl;
        }
        __syncthreads();
        const int ls_true = ls_run + s_eos[warp] + __popc(be & lt);
        const int ls = min(ls_true, MERGE_MAX_SEC - 1);
        const bool valid = in && !is_eos && (int)adr < p.npix_per_roach && sec0 + ls_true < p.exptime;
        if (!COPY) {
            if (valid) atomicAdd(&s_cnt[ls], 1);
        } else {
            const unsigned bv = __ballot_sync(0xffffffffu, valid);
            if (lane == 0) s_val[warp] = __popc(bv);
            __syncthreads();
            if (warp == 0) {
                const int v = s_val[lane];
                int incl = v;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) { const int a = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += a; }
                s_val[lane] = incl - v;
                if (lane == 31) s_val[32] = incl;
            }
            __syncthreads();
            if (valid) {
                const int rank = val_run + s_val[warp] + __popc(bv & lt);       // valid words of the segment before this one
                const long long pos = (long long)s_base[ls] + (rank - done_before[ls]);
                if (pos < p.out_cap) p.out[pos] = x;
            }
            val_run += s_val[32];
        }
        ls_run += s_eos[32];
        __syncthreads();
    }
    if (!COPY && tid < MERGE_MAX_SEC) p.cnt[tid * p.n_seg + seg] = s_cnt[tid];
}

// ------------------------------------------------------------------------------------------------
// Per-(second, pixel) photon lists: the product PacketMaster writes every second (photons[r][adr][plist],
// PacketMaster.c:371-380, write_sec_data :1012-1016).  Key = sec * n_pix + pixel, arrival order inside a key,
// each key truncated to max_events - 1 entries (the cap quirk).  Built from the rows of the relative pass:
//   list_rowstart_kernel  thread = (roach, channel) walks the roach's ranges in stream order: rows[r][ls][ch] is
//                         replaced by the number of earlier words of the same (second, pixel); the per-key totals
//                         land in acc[sec][pixel]
//   list_offsets_kernel   offsets = exclusive scan of min(acc, max_events - 1)
//   list_scatter_kernel   one CTA per range re-reads its words in order, sorts blocks of 2048 words by pixel in shared
//                         memory (stable) and stores word -> out[offsets[key] + rank] as whole sectors
struct ListParams {
    const uint64_t *words;
    const uint32_t *wire;         // wire format (PulseServer bundles) instead of words
    const DecRange *ranges;
    const DecRangeOut *rout;
    const int32_t *base, *eos_tot;
    const int32_t *roach_first;   // [n_roaches + 1] first range of every roach (ranges are sorted by roach)
    uint32_t *rows;
    uint32_t *acc;                // [exptime][n_pix], zeroed
    long long *offsets;           // [exptime * n_pix + 1]
    uint64_t *out;
    long long out_cap;
    int n_ranges, n_roaches, n_pix, npix_per_roach, exptime, cap;     // cap = max_events - 1
    int *flag;                    // set to 2 if out_cap is too small
};

// The (row, second) entries of the ranges [rb, rb + 256) of one roach in stream order, seconds >= exptime left out:
// built by the whole CTA (256 threads) so that the walks below are tight loops over shared memory
struct RowEntries {
    int row[256 * DEC_MAX_LS], sec[256 * DEC_MAX_LS];      // row = range * DEC_MAX_LS + local second
    int n, wsum[8];
};
__device__ __forceinline__ void build_row_entries(const ListParams &p, int rb, int r1, RowEntries &E) {
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5, r = rb + t;
    int nls = 0, base = 0;
    if (r < r1) { nls = p.rout[r].n_ls; base = p.base[r]; }
    const int cnt = max(0, min(nls, p.exptime - base));
    int incl = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const int a = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += a; }
    if (lane == 31) E.wsum[warp] = incl;
    __syncthreads();
    int before = 0;
#pragma unroll
    for (int q = 0; q < 8; ++q) before += q < warp ? E.wsum[q] : 0;
    int at = before + incl - cnt;
    for (int ls = 0; ls < cnt; ++ls, ++at) { E.row[at] = r * DEC_MAX_LS + ls; E.sec[at] = base + ls; }
    if (t == 255) E.n = before + incl;
    __syncthreads();
}

__global__ void __launch_bounds__(256) list_rowstart_kernel(ListParams p) {
    __shared__ RowEntries E;
    const int roach = blockIdx.x, ch = threadIdx.x;
    const int r0 = p.roach_first[roach], r1 = p.roach_first[roach + 1];
    const bool is_pix = ch < p.npix_per_roach;
    int cur_sec = -1, max_sec = -1;
    uint32_t run = 0;
    uint32_t *acc_col = p.acc + (size_t)roach * p.npix_per_roach + ch;
    uint32_t *col = p.rows + ch;
    for (int rb = r0; rb < r1; rb += 256) {
        build_row_entries(p, rb, r1, E);
        const int n = E.n;
        if (is_pix)
            for (int e0 = 0; e0 < n; e0 += 16) {   // 16 entries per round: their loads are independent of the running count
                uint32_t v[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = e0 + i < n ? col[(size_t)E.row[e0 + i] * DEC_ROW] : 0u;
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    if (e0 + i >= n) break;
                    const int sec = E.sec[e0 + i];
                    if (sec != cur_sec) {          // seconds normally only grow along a roach stream: rare
                        if (cur_sec >= 0) acc_col[(size_t)cur_sec * p.n_pix] = run;
                        cur_sec = sec;
                        // acc is zeroed before this kernel and this thread is the only writer of its column: a second
                        // that was not visited yet needs no (dependent, ~1 us) read
                        run = sec > max_sec ? 0u : acc_col[(size_t)sec * p.n_pix];
                        max_sec = max(max_sec, sec);
                    }
                    col[(size_t)E.row[e0 + i] * DEC_ROW] = run;
                    run += v[i];
                }
            }
        __syncthreads();
    }
    if (is_pix && cur_sec >= 0) acc_col[(size_t)cur_sec * p.n_pix] = run;
}

// offsets = exclusive scan of min(acc, cap) over the keys: block sums, scan of the block sums, block scans
constexpr int LIST_SCAN_BLOCK = 4096;                 // keys per CTA (1024 threads x 4, coalesced 128-bit loads)
__global__ void __launch_bounds__(1024) list_blocksum_kernel(ListParams p, long long *block_sums) {
    __shared__ long long s_w[32];
    const long long n = (long long)p.exptime * p.n_pix;
    const long long i = ((long long)blockIdx.x * 1024 + threadIdx.x) * 4;
    long long v = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) if (i + k < n) v += min(p.acc[i + k], (uint32_t)p.cap);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x < 32) {
        long long t = s_w[threadIdx.x];
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) t += __shfl_xor_sync(0xffffffffu, t, d);
        if (threadIdx.x == 0) block_sums[blockIdx.x] = t;
    }
}
__global__ void __launch_bounds__(1024) list_blockscan_kernel(ListParams p, long long *block_sums, int n_blocks) {
    __shared__ long long s_part[1024];          // exclusive scan of the block sums in place (one CTA)
    const int t = threadIdx.x;
    const int per = (n_blocks + 1023) / 1024, i0 = min(n_blocks, t * per), i1 = min(n_blocks, i0 + per);
    long long sum = 0;
    for (int i = i0; i < i1; ++i) sum += block_sums[i];
    s_part[t] = sum;
    __syncthreads();
    for (int d = 1; d < 1024; d <<= 1) {
        const long long a = t >= d ? s_part[t - d] : 0;
        __syncthreads();
        s_part[t] += a;
        __syncthreads();
    }
    long long run = t > 0 ? s_part[t - 1] : 0;
    for (int i = i0; i < i1; ++i) { const long long v = block_sums[i]; block_sums[i] = run; run += v; }
    if (t == 1023) {
        p.offsets[(long long)p.exptime * p.n_pix] = s_part[1023];
        if (s_part[1023] > p.out_cap) atomicOr(p.flag, 2);
    }
}
__global__ void __launch_bounds__(1024) list_offsets_kernel(ListParams p, const long long *block_sums) {
    __shared__ long long s_w[32];
    const long long n = (long long)p.exptime * p.n_pix;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const long long i = ((long long)blockIdx.x * 1024 + t) * 4;
    long long v[4], tot = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) { v[k] = i + k < n ? (long long)min(p.acc[i + k], (uint32_t)p.cap) : 0; tot += v[k]; }
    long long incl = tot;                       // inclusive scan over the warp, then over the 32 warps
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const long long a = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += a; }
    if (lane == 31) s_w[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        long long w = s_w[lane], wi = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const long long a = __shfl_up_sync(0xffffffffu, wi, d); if (lane >= d) wi += a; }
        s_w[lane] = wi - w;                     // exclusive
    }
    __syncthreads();
    long long run = block_sums[blockIdx.x] + s_w[warp] + incl - tot;
#pragma unroll
    for (int k = 0; k < 4; ++k) { if (i + k < n) p.offsets[i + k] = run; run += v[k]; }
}

// word `pos` of a range, host order; wire format: low and high halves sit in 4096-word blocks of big-endian u32
// (PulseServer.c:318-352), rg.start counts such blocks
template <bool WIRE>
__device__ __forceinline__ uint64_t list_load_word(const ListParams &p, const DecRange &rg, int pos, uint64_t pol_in) {
    if (WIRE) {
        const long long hc = rg.start + (pos >> 12);
        const uint32_t *lp = p.wire + (size_t)(hc >> 1) * (2 * DEC_BUNDLE) + (size_t)(hc & 1) * DEC_CHUNK + (pos & 4095);
        uint32_t lo, hi;
        asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.b32 %0, [%1], %2;" : "=r"(lo) : "l"(lp), "l"(pol_in));
        asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.b32 %0, [%1], %2;" : "=r"(hi) : "l"(lp + DEC_BUNDLE), "l"(pol_in));
        return (uint64_t)bswap32(hi) << 32 | bswap32(lo);
    } else {
        uint64_t v;
        asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.b64 %0, [%1], %2;" : "=l"(v) : "l"(p.words + rg.start + pos), "l"(pol_in));
        return v;
    }
}

// One CTA per range re-reads its words in blocks of 2048 and sorts every block by pixel in shared memory, stably
// (arrival order inside a pixel): warp w ranks words [256 w, 256 w + 256) of the block step by step (lane masks of
// the 32 words of a step + the warp's own running count per pixel), the counts of the warps are scanned per pixel and
// over the pixels, and the block leaves as one run per pixel.  Only whole, aligned 32-byte sectors are stored (single
// 8-byte stores scattered over the output cost a DRAM read-modify-write each: 2.9x traffic measured); the up to
// three words of a pixel that do not fill a sector yet are carried to the next block, partial sectors are written
// only at the start and the end of a (range, second).  A block ends early at an end-of-second word.
constexpr int LIST_WARPS = 8;                         // (also the warps of merge_scatter_kernel)
constexpr int LB_STEPS = 8, LB_BLOCK = LIST_WARPS * LB_STEPS * 32;
struct ListBlockSmem {
    uint64_t sorted[LB_BLOCK + 256 * 3];   // carried + new words of the block, pixel by pixel
    uint64_t carry[256][3];
    long long dst[256];                    // list index of the first carried (else next) word of every pixel
    uint32_t wmask[LIST_WARPS][256];       // lanes of the current step that hit pixel p (zero between steps)
    uint16_t wcnt[LIST_WARPS][256];        // words of pixel p ranked by warp w -> words of pixel p in earlier warps
    int left[256];                         // slots left under the cap
    uint32_t slot[256];                    // packed for the scatter: (take << 16) | (pstart + ncarry)
    long long outd[256];                   // packed for the copy-out: ((dst - pstart) << 12) | (pstart + nemit)
    unsigned char ncarry[256];
    uint32_t wsum[LIST_WARPS];
    int first_eos, total;
};
template <bool WIRE>
__global__ void __launch_bounds__(LIST_WARPS * 32) list_scatter_kernel(ListParams p) {
    __shared__ ListBlockSmem sm;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (*p.flag & 2) return;
    const unsigned lt = (1u << lane) - 1u;
    uint64_t pol_in;           // the input streams through L2 (evict first)
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_in));
    for (int i = tid; i < LIST_WARPS * 256; i += LIST_WARPS * 32) (&sm.wmask[0][0])[i] = 0u;
    for (int r = blockIdx.x; r < p.n_ranges; r += gridDim.x) {
        const DecRange rg = p.ranges[r];
        const int n_words = rg.n_words, base = p.base[r], n_ls = p.rout[r].n_ls, npix = p.npix_per_roach;
        int ls = 0, pos = 0;
        auto open_second = [&]() {              // thread = pixel: slots of local second ls (nothing is stored beyond exptime)
            const int sec = base + ls;
            long long d = 0; int l = 0;
            if (sec < p.exptime && ls < n_ls && tid < npix) {
                const uint32_t before = p.rows[((size_t)r * DEC_MAX_LS + ls) * DEC_ROW + tid];
                d = p.offsets[(long long)sec * p.n_pix + (long long)rg.roach * npix + tid] + before;
                l = before < (uint32_t)p.cap ? p.cap - (int)before : 0;
            }
            sm.dst[tid] = d; sm.left[tid] = l; sm.ncarry[tid] = 0;
        };
        auto flush_carry = [&]() {              // thread = pixel: what is carried but not yet a whole sector
            const int c = sm.ncarry[tid];
            const long long d = sm.dst[tid];
            for (int j = 0; j < c; ++j) if (d + j < p.out_cap) p.out[d + j] = sm.carry[tid][j];
            sm.ncarry[tid] = 0;
        };
        open_second();
        __syncthreads();
        while (pos < n_words) {
            const int rem = n_words - pos;
            uint64_t x[LB_STEPS];
#pragma unroll
            for (int s = 0; s < LB_STEPS; ++s) {
                const int idx = warp * (LB_STEPS * 32) + s * 32 + lane;
                x[s] = idx < rem ? list_load_word<WIRE>(p, rg, pos + idx, pol_in) : 0ull;
            }
            for (int i = tid; i < LIST_WARPS * 256 / 2; i += LIST_WARPS * 32) reinterpret_cast<uint32_t *>(&sm.wcnt[0][0])[i] = 0u;
            if (tid == 0) sm.first_eos = LB_BLOCK;
            __syncthreads();
            unsigned fe = LB_BLOCK;
#pragma unroll
            for (int s = LB_STEPS - 1; s >= 0; --s) {
                const int idx = warp * (LB_STEPS * 32) + s * 32 + lane;
                if (idx < rem && (uint32_t)(x[s] >> 56) == 255u) fe = idx;
            }
            fe = __reduce_min_sync(0xffffffffu, fe);
            if (lane == 0 && fe < (unsigned)LB_BLOCK) atomicMin(&sm.first_eos, (int)fe);
            __syncthreads();
            const int first_eos = sm.first_eos;
            const int n_here = min(first_eos, min(rem, LB_BLOCK));
            // rank of every word among the words of its pixel that this warp holds
            uint32_t rk[LB_STEPS];
            unsigned stored = 0;
#pragma unroll
            for (int s = 0; s < LB_STEPS; ++s) {
                const int idx = warp * (LB_STEPS * 32) + s * 32 + lane;
                const uint32_t adr = (uint32_t)(x[s] >> 56);
                const bool store = idx < n_here && (int)adr < npix;
                // lanes of this step that hit the same pixel, through shared memory: every lane sets its bit in the warp's
                // mask of the pixel and reads the mask back (MATCH.ANY and VOTE both run on the ADU pipe - about 64 and 6
                // cycles per SM each - and bounded this kernel: 95 % / 76 % ADU busy measured)
                if (store) atomicOr(&sm.wmask[warp][adr], 1u << lane);
                __syncwarp();
                uint32_t peers = 0, old = 0;
                if (store) { peers = sm.wmask[warp][adr]; old = sm.wcnt[warp][adr]; }
                __syncwarp();
                rk[s] = 0;
                if (store) {
                    if ((peers & lt) == 0u) { sm.wcnt[warp][adr] = (uint16_t)(old + __popc(peers)); sm.wmask[warp][adr] = 0u; }
                    rk[s] = old + __popc(peers & lt);
                    stored |= 1u << s;
                }
                __syncwarp();
            }
            __syncthreads();
            uint32_t my_n, my_ne;   // thread = pixel: scan over the warps, cap, scan over the pixels, sectors that can leave
            {
                uint32_t tot = 0;
#pragma unroll
                for (int q = 0; q < LIST_WARPS; ++q) { const uint32_t t = sm.wcnt[q][tid]; sm.wcnt[q][tid] = (uint16_t)tot; tot += t; }
                const int l = sm.left[tid];
                const uint32_t tk = min(tot, (uint32_t)l);
                sm.left[tid] = l - (int)tk;
                const uint32_t c = sm.ncarry[tid], n = c + tk;
                uint32_t incl = n;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) { const uint32_t a = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += a; }
                if (lane == 31) sm.wsum[warp] = incl;
                __syncthreads();
                uint32_t before = 0;
#pragma unroll
                for (int q = 0; q < LIST_WARPS; ++q) before += q < warp ? sm.wsum[q] : 0u;
                const uint32_t ps = before + incl - n;
                const long long d0 = sm.dst[tid], e_al = (d0 + n) & ~3ll;
                const uint32_t ne = (uint32_t)(e_al > d0 ? e_al - d0 : 0);
                my_n = n; my_ne = ne;
                if (tid == 255) sm.total = (int)(ps + n);
                sm.slot[tid] = (tk << 16) | (ps + c);
                sm.outd[tid] = ((d0 - (long long)ps) << 12) | (long long)(ps + ne);
                for (uint32_t j = 0; j < c; ++j) sm.sorted[ps + j] = sm.carry[tid][j];
            }
            __syncthreads();
#pragma unroll
            for (int s = 0; s < LB_STEPS; ++s) {
                if ((stored >> s) & 1u) {
                    const uint32_t adr = (uint32_t)(x[s] >> 56);
                    const uint32_t k = sm.wcnt[warp][adr] + rk[s], sl = sm.slot[adr];
                    if (k < (sl >> 16)) sm.sorted[(sl & 0xFFFFu) + k] = x[s];
                }
            }
            __syncthreads();
            const int total = sm.total;
            for (int i = tid; i < total; i += LIST_WARPS * 32) {
                const uint64_t v = sm.sorted[i];
                const uint32_t adr = (uint32_t)(v >> 56);
                const long long od = sm.outd[adr];
                const int e_end = (int)(od & 0xFFF);             // first slot of the pixel's run that stays behind
                if (i < e_end) {
                    const long long at = (od >> 12) + i;
                    if (at < p.out_cap) p.out[at] = v;
                } else {
                    sm.carry[adr][i - e_end] = v;
                }
            }
            __syncthreads();
            sm.dst[tid] += my_ne; sm.ncarry[tid] = (unsigned char)(my_n - my_ne);
            pos += n_here;
            if (first_eos < LB_BLOCK) {         // the end-of-second word at pos closes the local second
                flush_carry();
                ++ls; ++pos;
                open_second();
            }
            __syncthreads();
        }
        flush_carry();
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// Time-ordered merged photon list ("merged photon list sorted by (sec, roach, ts)"): every valid
// pixel word of seconds < exptime, key = sec * n_roaches + roach, stream order inside a key (the timestamps of one
// board only grow inside a second).  No cap: that is a property of the per-pixel lists.  Built from the same rows:
//   merge_rangesum_kernel   warp per range: rsum[r][ls] = words of local second ls that hit a pixel
//   merge_rowstart_kernel   CTA per roach: rsum[r][ls] <- words of earlier ranges with the same key; acc[key] = total
//   list_block*/offsets     offsets = exclusive scan of acc (shared with the per-pixel lists, n_pix := n_roaches)
//   merge_scatter_kernel    warp per range re-reads its words in order and stores the pixel words of every 32-word step
//                           as one contiguous run (ballot compaction): coalesced, sectors merge in L2
__global__ void __launch_bounds__(256) merge_rangesum_kernel(ListParams p, uint32_t *rsum) {
    const int r = (blockIdx.x * 256 + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (r >= p.n_ranges) return;
    const int n_ls = p.rout[r].n_ls;
#pragma unroll
    for (int ls = 0; ls < DEC_MAX_LS; ++ls) {
        uint32_t v = 0;
        if (ls < n_ls) {
            const uint32_t *row = p.rows + ((size_t)r * DEC_MAX_LS + ls) * DEC_ROW;
            for (int i = lane; i < p.npix_per_roach; i += 32) v += row[i];
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
        if (lane == 0) rsum[(size_t)r * DEC_MAX_LS + ls] = v;
    }
}

__global__ void __launch_bounds__(256) merge_rowstart_kernel(ListParams p, uint32_t *rsum) {
    __shared__ RowEntries E;
    __shared__ uint32_t s_v[256 * DEC_MAX_LS];
    const int roach = blockIdx.x, t = threadIdx.x;
    const int r0 = p.roach_first[roach], r1 = p.roach_first[roach + 1];
    for (int rb = r0; rb < r1; rb += 256) {
        build_row_entries(p, rb, r1, E);
        const int n = E.n;
        for (int e = t; e < n; e += 256) s_v[e] = rsum[E.row[e]];
        __syncthreads();
        // every entry sums the earlier entries of its second (at most 1024 entries: brute force over shared memory, all
        // threads read the same entry at a time); acc carries the seconds from chunk to chunk
        for (int e = t; e < n; e += 256) {
            const int sec = E.sec[e];
            uint32_t sum = p.acc[(size_t)sec * p.n_roaches + roach];
            bool last = true;
            for (int q = 0; q < n; ++q) {
                const bool same = E.sec[q] == sec;
                if (same && q < e) sum += s_v[q];
                if (same && q > e) last = false;
            }
            rsum[E.row[e]] = sum;
            if (last) p.acc[(size_t)sec * p.n_roaches + roach] = sum + s_v[e];
        }
        __syncthreads();
    }
}

template <bool WIRE>
__global__ void __launch_bounds__(LIST_WARPS * 32) merge_scatter_kernel(ListParams p, const uint32_t *rsum) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (*p.flag & 2) return;
    const unsigned lt = (1u << lane) - 1u;
    uint64_t pol_in;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_in));
    for (int r = blockIdx.x * LIST_WARPS + warp; r < p.n_ranges; r += gridDim.x * LIST_WARPS) {
        const DecRange rg = p.ranges[r];
        const int n_words = rg.n_words, base = p.base[r], n_ls = p.rout[r].n_ls, npix = p.npix_per_roach;
        int ls = 0;
        auto open_second = [&]() -> long long {        // next output index of local second ls, -1: nothing is stored
            const int sec = base + ls;
            if (sec >= p.exptime || ls >= n_ls) return -1ll;
            return p.offsets[(long long)sec * p.n_roaches + rg.roach] + rsum[(size_t)r * DEC_MAX_LS + ls];
        };
        long long dst = open_second();
        constexpr int LR = 8;
        uint64_t ring[LR];
#pragma unroll
        for (int g = 0; g < LR; ++g) ring[g] = g * 32 + lane < n_words ? list_load_word<WIRE>(p, rg, g * 32 + lane, pol_in) : 0ull;
        for (int pos0 = 0; pos0 < n_words; pos0 += 32 * LR) {
#pragma unroll
            for (int g = 0; g < LR; ++g) {
                const int pos = pos0 + g * 32;
                if (pos >= n_words) break;
                const uint64_t x = ring[g];
                if (pos + 32 * LR + lane < n_words) ring[g] = list_load_word<WIRE>(p, rg, pos + 32 * LR + lane, pol_in);
                const bool valid = pos + lane < n_words;
                const uint32_t adr = (uint32_t)(x >> 56);
                unsigned eos = __ballot_sync(0xffffffffu, valid && adr == 255u);
                unsigned todo = __ballot_sync(0xffffffffu, valid);
                while (todo) {
                    const int e = eos ? __ffs(eos) - 1 : 32;
                    const unsigned seg = todo & (e == 32 ? 0xFFFFFFFFu : ((1u << e) - 1u));
                    const bool store = ((seg >> lane) & 1u) && (int)adr < npix && dst >= 0;
                    const unsigned m = __ballot_sync(0xffffffffu, store);
                    if (store) {
                        const long long at = dst + __popc(m & lt);
                        if (at < p.out_cap) p.out[at] = x;
                    }
                    if (dst >= 0) dst += __popc(m);
                    todo &= ~seg;
                    if (e < 32) {
                        todo &= ~(1u << e);
                        eos &= ~(1u << e);
                        ++ls;
                        dst = open_second();
                    }
                }
            }
        }
    }
}

__global__ void counts_cap_kernel(const uint32_t *in, uint32_t *out, int64_t n, uint32_t cap) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) out[i] = min(in[i], cap);
}

__global__ void unpack_fields_kernel(const uint64_t *__restrict__ w, int64_t n, uint8_t *ch, uint32_t *ts,
                                     uint16_t *base, uint16_t *peak, uint16_t *p1) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        const uint64_t x = w[i];
        const uint32_t hi = (uint32_t)(x >> 32), lo = (uint32_t)x;
        if (ch) ch[i] = (uint8_t)(hi >> 24);
        if (peak) peak[i] = (uint16_t)((hi >> 12) & 0xFFF);
        if (p1) p1[i] = (uint16_t)(hi & 0xFFF);
        if (base) base[i] = (uint16_t)((lo >> 20) & 0xFFF);
        if (ts) ts[i] = lo & 0xFFFFF;
    }
}

__global__ void reinterpret_bin_kernel(const uint64_t *__restrict__ v, int64_t n, int n_bits, int binary_point,
                                       int after, double *out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const uint64_t mask = n_bits >= 64 ? ~0ull : ((1ull << n_bits) - 1);
    const double scale = exp2(-(double)binary_point);   // exact power of two
    for (; i < n; i += stride) {
        uint64_t x = (v[i] >> after) & mask;
        double d;
        if ((x >> (n_bits - 1)) & 1ull) d = -(double)(((~x) & mask) + 1ull);
        else d = (double)x;
        out[i] = d * scale;
    }
}

__global__ void quicklook_kernel(const uint32_t *counts_sec, const int32_t *pixel_adr, int n, uint16_t *image) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) image[i] = (uint16_t)counts_sec[pixel_adr[i]];
}

struct ListRequest { uint64_t *list_words; int64_t list_cap; int64_t *list_offsets; bool by_roach; };

// persistent device buffer i of the decode path (contents survive other calls on the context; lost on growth)
int dec_private(mkid_ctx *ctx, int i, size_t bytes, void **out) {
    if (ctx->dec_priv_bytes[i] < bytes) {
        MKID_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (ctx->dec_priv[i]) cudaFree(ctx->dec_priv[i]);
        ctx->dec_priv[i] = nullptr; ctx->dec_priv_bytes[i] = 0;
        const size_t cap = (bytes + 4095) / 4096 * 4096 * 2;
        if (cudaMalloc(&ctx->dec_priv[i], cap) != cudaSuccess) return mkid_fail(ctx, MKID_ENOMEM, "cudaMalloc(%zu) failed", cap);
        ctx->dec_priv_bytes[i] = cap;
        if (i == 0) ctx->dec_meta_dev = nullptr; else ctx->dec_ranges_dev = nullptr;
    }
    *out = ctx->dec_priv[i];
    return MKID_OK;
}

int decode_common(mkid_ctx *ctx, const uint64_t *words, const uint32_t *wire, int64_t n_units,
                  const int64_t *seg_offset, const int64_t *seg_len_in, const int32_t *seg_roach, const int32_t *seg_sec,
                  int32_t *seg_sec_out, int32_t n_seg, const mkid_decode_cfg *cfg, uint32_t *counts_raw,
                  uint32_t *hist, mkid_decode_stats *stats, const int32_t *seg_len_dev = nullptr,
                  const int32_t *seg_sec_dev = nullptr, int32_t *seg_sec_out_dev = nullptr, const ListRequest *lists = nullptr) {
    MKID_REQUIRE(ctx, cfg && seg_offset && seg_roach && n_seg > 0, "decode: missing cfg/segments");
    MKID_REQUIRE(ctx, cfg->npix_per_roach > 0 && cfg->npix_per_roach <= 255, "npix_per_roach must be 1..255");
    MKID_REQUIRE(ctx, cfg->n_roaches > 0 && cfg->exptime > 0 && counts_raw, "bad decode cfg");
    const bool want_hist = hist != nullptr && cfg->hist_field_shift >= 0;
    if (want_hist) MKID_REQUIRE(ctx, cfg->n_bins > 0 && cfg->hist_field_shift <= 52, "bad histogram cfg");
    MKID_CUDA(ctx, cudaSetDevice(ctx->device));
    const bool wire_fmt = wire != nullptr;
    const int64_t n_pix = (int64_t)cfg->n_roaches * cfg->npix_per_roach;

    // segments -> units -> ranges: about one range per resident warp, whole units (flat: 256 words, wire: one
    // 4096-word half bundle), never across a segment, all ranges of (almost) equal length
    const int unit_words = wire_fmt ? DEC_CHUNK : DEC_UNIT;
    const int64_t min_units = wire_fmt ? 1 : 1024 / DEC_UNIT;     // no range shorter than 1024 words (wire: one chunk)
    std::vector<int64_t> seg_chunks(n_seg), seg_len(n_seg);     // seg_chunks: units per segment
    int64_t n_chunks = 0;
    for (int i = 0; i < n_seg; ++i) {
        const int64_t len = seg_len_in ? seg_len_in[i] : seg_offset[i + 1] - seg_offset[i];
        seg_len[i] = len;
        MKID_REQUIRE(ctx, len >= 0 && seg_offset[i] >= 0 && seg_offset[i] + len <= n_units, "segment offsets out of range");
        MKID_REQUIRE(ctx, seg_roach[i] >= 0 && seg_roach[i] < cfg->n_roaches, "segment roach out of range");
        seg_chunks[i] = wire_fmt ? 2 * len : (len + DEC_UNIT - 1) / DEC_UNIT;
        n_chunks += seg_chunks[i];
    }
    std::vector<int32_t> sec0(n_seg, 0);
    if (seg_sec) for (int i = 0; i < n_seg; ++i) sec0[i] = seg_sec[i];

    // short segments behind a producer on the same GPU: one launch (decode_small_kernel), a segment table instead of ranges
    const bool small = seg_len_dev && !wire_fmt && !lists && n_units <= DEC_SMALL_MAX_WORDS && n_seg <= 65535;
    // the range table only depends on the segment table: rebuilt (and uploaded) when that changes
    std::vector<char> key(24 + (size_t)n_seg * 20);
    {
        const int64_t head[3] = {wire_fmt ? 1 : (small ? 2 : 0), n_seg, n_units};
        memcpy(key.data(), head, 24);
        memcpy(key.data() + 24, seg_offset, (size_t)n_seg * 8);
        memcpy(key.data() + 24 + (size_t)n_seg * 8, seg_len.data(), (size_t)n_seg * 8);
        memcpy(key.data() + 24 + (size_t)n_seg * 16, seg_roach, (size_t)n_seg * 4);
    }
    const bool same_table = key == ctx->dec_key;
    std::vector<DecRange> ranges;
    if (!same_table && small) {
        std::vector<DecSeg> st(n_seg);
        for (int i = 0; i < n_seg; ++i) {
            MKID_REQUIRE(ctx, seg_len[i] <= 0x7FFFFFFF, "decode: segment too long");
            st[i].start = seg_offset[i]; st[i].cap = (int)seg_len[i]; st[i].roach = seg_roach[i];
        }
        ctx->dec_ranges_host.assign((const char *)st.data(), (const char *)st.data() + st.size() * sizeof(DecSeg));
        ctx->dec_ranges_dev = nullptr;
        ctx->dec_key.swap(key);
    } else if (!same_table) {
        const int64_t warps_total = (int64_t)ctx->num_sms * DEC_CTAS_PER_SM * DEC_WARPS;
        for (int i = 0; i < n_seg; ++i) {
            const int64_t nc = seg_chunks[i];
            if (nc == 0) continue;
            int64_t pieces = (int64_t)((double)nc * (double)warps_total / (double)n_chunks);        // floor: total <= warps_total
            pieces = std::max<int64_t>(1, std::min<int64_t>(pieces, nc / min_units));
            for (int64_t q = 0; q < pieces; ++q) {
                const int64_t c0 = nc * q / pieces, c1 = nc * (q + 1) / pieces;
                if (c1 <= c0) continue;
                DecRange r;
                if (wire_fmt) { r.start = seg_offset[i] * 2 + c0; r.n_words = (int)((c1 - c0) * DEC_CHUNK); }
                else { r.start = seg_offset[i] + c0 * DEC_UNIT; r.n_words = (int)std::min<int64_t>((c1 - c0) * DEC_UNIT, seg_len[i] - c0 * DEC_UNIT); }
                r.roach = seg_roach[i]; r.seg = i; r.seg_first = 0; r.pad = 0;
                r.seg_off = (int)(c0 * unit_words);
                ranges.push_back(r);
            }
        }
        // ranges of one roach next to each other (segments stay contiguous and in order): the 16 warps of a CTA then
        // share one shared-memory histogram
        std::stable_sort(ranges.begin(), ranges.end(), [](const DecRange &a, const DecRange &b) { return a.roach < b.roach; });
        for (int k = 0; k < (int)ranges.size(); ++k) ranges[k].seg_first = (k > 0 && ranges[k - 1].seg == ranges[k].seg) ? ranges[k - 1].seg_first : k;
        MKID_REQUIRE(ctx, ranges.size() <= (size_t)1 << 18, "decode: too many segments in one call");
        for (const DecRange &r : ranges) MKID_REQUIRE(ctx, (int64_t)r.n_words <= (int64_t)1 << 30, "decode: range too long");

        ctx->dec_ranges_host.assign((const char *)ranges.data(), (const char *)ranges.data() + ranges.size() * sizeof(DecRange));
        ctx->dec_ranges_dev = nullptr;          // forces the upload below
        ctx->dec_key.swap(key);
    }
    const int n_ranges = small ? 0 : (int)(ctx->dec_ranges_host.size() / sizeof(DecRange));
    // meta: stats (5 u64) | flag, pad | sec_out [n_seg] | sec [n_seg]
    const size_t meta_bytes = 48 + (size_t)n_seg * 8;
    char *meta = nullptr;
    int rc = dec_private(ctx, 0, meta_bytes, (void **)&meta);
    if (rc) return rc;
    unsigned long long *d_stats = (unsigned long long *)meta;
    int *d_flag = (int *)(meta + 40);
    int32_t *d_sec_out = (int32_t *)(meta + 48);
    int32_t *d_sec = d_sec_out + n_seg;
    {   // upload the segment seconds only when they changed since the last call on this context
        std::vector<char> blob(meta_bytes, 0);
        memcpy(blob.data() + 48 + (size_t)n_seg * 4, sec0.data(), (size_t)n_seg * 4);
        if (ctx->dec_meta_dev != meta || ctx->dec_meta_host != blob) {
            MKID_CUDA(ctx, cudaMemcpyAsync(meta, blob.data(), meta_bytes, cudaMemcpyHostToDevice, ctx->stream));
            MKID_CUDA(ctx, cudaStreamSynchronize(ctx->stream));   // blob is pageable and about to be moved
            ctx->dec_meta_host.swap(blob);
            ctx->dec_meta_dev = meta;
        } else {
            MKID_CUDA(ctx, cudaMemsetAsync(meta, 0, 48 + (size_t)n_seg * 4, ctx->stream));
        }
    }

    const void *d_in = nullptr;
    const size_t in_bytes = wire_fmt ? (size_t)n_units * 2 * DEC_BUNDLE * 4 : (size_t)n_units * 8;
    rc = mkid_stage_in(ctx, wire_fmt ? (const void *)wire : (const void *)words, in_bytes, SCR_IN, &d_in);
    if (rc) return rc;
    MKID_REQUIRE(ctx, (reinterpret_cast<uintptr_t>(d_in) & (wire_fmt ? 15 : 7)) == 0, "decode: input must be 8-byte (words) / 16-byte (wire) aligned");
    const void *d_lut = nullptr;
    if (want_hist && cfg->bin_lut) {
        rc = mkid_stage_in(ctx, cfg->bin_lut, 4096 * 2, SCR_AUX0, &d_lut);
        if (rc) return rc;
    }
    void *d_counts = nullptr, *d_hist = nullptr;
    const size_t counts_bytes = (size_t)cfg->exptime * n_pix * 4;
    const size_t hist_bytes = want_hist ? (size_t)n_pix * cfg->n_bins * 4 : 0;
    rc = mkid_stage_out(ctx, counts_raw, counts_bytes, SCR_OUT0, true, &d_counts);
    if (rc) return rc;
    if (want_hist) {
        rc = mkid_stage_out(ctx, hist, hist_bytes, SCR_OUT1, true, &d_hist);
        if (rc) return rc;
    }

    if (small) {
        DecSeg *d_segs;
        if ((rc = dec_private(ctx, 1, ctx->dec_ranges_host.size(), (void **)&d_segs))) return rc;
        if (ctx->dec_ranges_dev != d_segs) {
            MKID_CUDA(ctx, cudaMemcpyAsync(d_segs, ctx->dec_ranges_host.data(), ctx->dec_ranges_host.size(), cudaMemcpyHostToDevice, ctx->stream));
            MKID_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            ctx->dec_ranges_dev = d_segs;
        }
        DecParams p;
        memset(&p, 0, sizeof(p));
        p.words = (const uint64_t *)d_in;
        p.seg_sec = seg_sec_dev ? seg_sec_dev : d_sec; p.seg_sec_out = seg_sec_out_dev ? seg_sec_out_dev : d_sec_out;
        p.seg_len_dev = seg_len_dev;
        p.n_pix = (int)n_pix; p.npix_per_roach = cfg->npix_per_roach; p.exptime = cfg->exptime;
        p.field_shift = want_hist ? cfg->hist_field_shift : 0; p.n_bins = want_hist ? cfg->n_bins : 0;
        p.bin_lut = (const uint16_t *)d_lut; p.counts = (uint32_t *)d_counts; p.hist = (uint32_t *)d_hist;
        p.stats = d_stats;
        decode_small_kernel<<<n_seg, 1024, 0, ctx->stream>>>(p, d_segs);
        MKID_CHECK_LAUNCH(ctx);
    }
    if (n_ranges > 0) {
        DecRange *d_ranges; DecRangeOut *d_rout; uint32_t *d_rows;
        if ((rc = dec_private(ctx, 1, (size_t)n_ranges * sizeof(DecRange), (void **)&d_ranges))) return rc;
        if ((rc = mkid_scratch(ctx, SCR_STATE, (size_t)n_ranges * (sizeof(DecRangeOut) + 8), (void **)&d_rout))) return rc;
        if ((rc = mkid_scratch(ctx, SCR_AUX4, (size_t)n_ranges * DEC_MAX_LS * DEC_ROW * 4, (void **)&d_rows))) return rc;
        int32_t *d_base = (int32_t *)(d_rout + n_ranges);
        int32_t *d_eos = d_base + n_ranges;
        if (ctx->dec_ranges_dev != d_ranges) {
            MKID_CUDA(ctx, cudaMemcpyAsync(d_ranges, ctx->dec_ranges_host.data(), ctx->dec_ranges_host.size(), cudaMemcpyHostToDevice, ctx->stream));
            MKID_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            ctx->dec_ranges_dev = d_ranges;
        }
        DecParams p;
        p.words = wire_fmt ? nullptr : (const uint64_t *)d_in;
        p.wire = wire_fmt ? (const uint32_t *)d_in : nullptr;
        p.ranges = d_ranges; p.rout = d_rout; p.n_ranges = n_ranges; p.rows = d_rows; p.base = d_base; p.flag = d_flag;
        p.eos_tot = d_eos; p.seg_sec = seg_sec_dev ? seg_sec_dev : d_sec; p.seg_sec_out = seg_sec_out_dev ? seg_sec_out_dev : d_sec_out;
        p.seg_len_dev = seg_len_dev;
        p.n_pix = (int)n_pix; p.npix_per_roach = cfg->npix_per_roach; p.exptime = cfg->exptime;
        p.field_shift = want_hist ? cfg->hist_field_shift : 0; p.n_bins = want_hist ? cfg->n_bins : 0;
        p.bin_lut = (const uint16_t *)d_lut; p.counts = (uint32_t *)d_counts; p.hist = (uint32_t *)d_hist;
        p.stats = d_stats;

        const bool smem_hist = want_hist && (int64_t)cfg->npix_per_roach * cfg->n_bins <= DEC_SMEM_HIST;
        const int grid = (int)std::min<int64_t>((n_ranges + DEC_WARPS - 1) / DEC_WARPS, (int64_t)ctx->num_sms * DEC_CTAS_PER_SM);
        const bool need_lo = want_hist && cfg->hist_field_shift < 32;
        // A histogram too large for shared memory can be built from partitioned 16-bit keys (HIST == 3, PartParams) instead
        // of in place: MKID_DEC_PART=1.  Bit-identical (tests/test_decode_gpu.py::test_partitioned_histogram_forms), but off
        // by default: on 1.6e8 words x [2024][4096] the key pass takes 0.69 ms (86 instructions per word, issue- and
        // scoreboard-bound) + 0.21 ms for the tiles against 0.97 ms in place.
        bool part = false;
        p.part = PartParams{nullptr, nullptr, 0u, 0, 0};
        if (want_hist && !smem_hist && cfg->n_bins <= 4096) {
            const int64_t unit = wire_fmt ? DEC_BUNDLE : 1;
            std::vector<int64_t> w_roach(cfg->n_roaches, 0), r_roach(cfg->n_roaches, 0);
            for (int i = 0; i < n_seg; ++i) w_roach[seg_roach[i]] += seg_len[i] * unit;
            const DecRange *hr = (const DecRange *)ctx->dec_ranges_host.data();
            for (int k = 0; k < n_ranges; ++k) r_roach[hr[k].roach]++;
            const char *e = getenv("MKID_DEC_PART");
            part = e && atoi(e) != 0;
            const int tp = (cfg->npix_per_roach + PART_TILES - 1) / PART_TILES;                // <= 8 pixels per tile
            const int n_tiles = (cfg->npix_per_roach + tp - 1) / tp;
            // a bucket holds four times the tile's share of the roach's words (a tile that takes more sends the surplus to
            // the histogram directly) plus what the ranges of the roach can leave unused in their last chunks
            int64_t cap = 0;
            for (int q = 0; q < cfg->n_roaches; ++q) {
                const int64_t share = n_tiles <= 4 ? w_roach[q] : std::min<int64_t>(w_roach[q], 4 * ((w_roach[q] + n_tiles - 1) / n_tiles));
                cap = std::max<int64_t>(cap, share + r_roach[q] * (2 * PART_CHUNK + PART_FLUSH));
            }
            cap = (cap + PART_CHUNK - 1) / PART_CHUNK * PART_CHUNK + PART_CHUNK;
            if (cap >= ((int64_t)1 << 31)) part = false;
            if (part) {
                const size_t n_buckets = (size_t)cfg->n_roaches * n_tiles;
                char *pbuf;
                if ((rc = mkid_scratch(ctx, SCR_AUX1, n_buckets * 4 + 256 + n_buckets * (size_t)cap * 2, (void **)&pbuf))) return rc;
                p.part.cursor = (uint32_t *)pbuf;
                p.part.keys = (uint16_t *)(pbuf + (n_buckets * 4 + 255) / 256 * 256);
                p.part.cap = (uint32_t)cap; p.part.n_tiles = n_tiles; p.part.tp = tp;
                MKID_CUDA(ctx, cudaMemsetAsync(p.part.cursor, 0, n_buckets * 4, ctx->stream));
            }
        }
        const int hist_mode = !want_hist ? 0 : smem_hist ? 2 : part ? 3 : 1;
        constexpr size_t part_smem = (size_t)DEC_WARPS * PART_SMEM_WARP;
        auto launch = [&](auto kern) { kern<<<grid, DEC_THREADS, 0, ctx->stream>>>(p); };
        auto launch_part = [&](auto kern) {
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)part_smem);
            kern<<<grid, DEC_THREADS, part_smem, ctx->stream>>>(p);
        };
        auto stream_pass = [&](auto abs_tag) {
            constexpr bool ABS = decltype(abs_tag)::value;       // (the absolute pass is rare: one variant per mode)
            constexpr bool LO_ALWAYS = ABS;
            if (!ABS && hist_mode == 3) {                        // (the absolute pass of the partitioned form reduces in place)
                if (wire_fmt) { if (need_lo) launch_part(decode_stream_kernel<true, 3, true, false>); else launch_part(decode_stream_kernel<true, 3, false, false>); }
                else { if (need_lo) launch_part(decode_stream_kernel<false, 3, true, false>); else launch_part(decode_stream_kernel<false, 3, false, false>); }
                return;
            }
            if (wire_fmt) {
                if (hist_mode == 0) launch(decode_stream_kernel<true, 0, LO_ALWAYS, ABS>);
                else if (hist_mode == 1 || hist_mode == 3) { if (need_lo || LO_ALWAYS) launch(decode_stream_kernel<true, 1, true, ABS>); else launch(decode_stream_kernel<true, 1, LO_ALWAYS, ABS>); }
                else { if (need_lo || LO_ALWAYS) launch(decode_stream_kernel<true, 2, true, ABS>); else launch(decode_stream_kernel<true, 2, LO_ALWAYS, ABS>); }
            } else {
                if (hist_mode == 0) launch(decode_stream_kernel<false, 0, LO_ALWAYS, ABS>);
                else if (hist_mode == 1 || hist_mode == 3) { if (need_lo || LO_ALWAYS) launch(decode_stream_kernel<false, 1, true, ABS>); else launch(decode_stream_kernel<false, 1, LO_ALWAYS, ABS>); }
                else { if (need_lo || LO_ALWAYS) launch(decode_stream_kernel<false, 2, true, ABS>); else launch(decode_stream_kernel<false, 2, LO_ALWAYS, ABS>); }
            }
        };
        // MKID_DEC_TIMING=1: per-kernel device times on stderr (CUDA events; synchronises)
        static const bool timing = getenv("MKID_DEC_TIMING") != nullptr;
        cudaEvent_t *ev = ctx->dbg_events;
        auto mark = [&](int i) {
            if (!timing) return;
            if (!ev[i]) cudaEventCreate(&ev[i]);
            cudaEventRecord(ev[i], ctx->stream);
        };
        mark(0);
        stream_pass(std::false_type{});                      // relative pass: every word is read here, once
        MKID_CHECK_LAUNCH(ctx);
        static cudaEvent_t ev_part = nullptr;            // (timing switch only)
        if (hist_mode == 3) {
            if (timing) { if (!ev_part) cudaEventCreate(&ev_part); cudaEventRecord(ev_part, ctx->stream); }
            const size_t hsm = (size_t)p.part.tp * cfg->n_bins * 4;
            MKID_CUDA(ctx, cudaFuncSetAttribute(part_hist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hsm));
            part_hist_kernel<<<cfg->n_roaches * p.part.n_tiles, 1024, hsm, ctx->stream>>>(p.part, p.hist, cfg->npix_per_roach, cfg->n_bins);
            MKID_CHECK_LAUNCH(ctx);
        }
        mark(1);
        const int cgrid = (n_ranges + DEC_COMMIT_RANGES - 1) / DEC_COMMIT_RANGES;
        if (wire_fmt) decode_commit_kernel<true><<<cgrid, 256, 0, ctx->stream>>>(p);
        else decode_commit_kernel<false><<<cgrid, 256, 0, ctx->stream>>>(p);
        MKID_CHECK_LAUNCH(ctx);
        mark(2);
        stream_pass(std::true_type{});                       // absolute pass: exits at once unless a range overflowed
        MKID_CHECK_LAUNCH(ctx);
        mark(3);
        if (timing) {
            float t[3];
            MKID_CUDA(ctx, cudaEventSynchronize(ev[3]));
            for (int i = 0; i < 3; ++i) cudaEventElapsedTime(&t[i], ev[i], ev[i + 1]);
            fprintf(stderr, "[mkid decode timing] ranges %d | stream %.1f us  commit %.1f us  absolute %.1f us\n", n_ranges,
                    t[0] * 1e3f, t[1] * 1e3f, t[2] * 1e3f);
            if (hist_mode == 3) {
                float tp1 = 0.f;
                cudaEventElapsedTime(&tp1, ev[0], ev_part);
                fprintf(stderr, "[mkid decode timing] partitioned histogram: keys %.1f us, tiles %.1f us (%d buckets of <= %u keys)\n", tp1 * 1e3f,
                        (t[0] - tp1) * 1e3f, cfg->n_roaches * p.part.n_tiles, p.part.cap);
            }
        }
        if (lists) {
            // the list product needs every range resolved by its rows: more than DEC_MAX_LS seconds inside one range
            // (>= 1024 words) is not supported here
            int flag_h = 0;
            MKID_CUDA(ctx, cudaMemcpyAsync(&flag_h, d_flag, 4, cudaMemcpyDeviceToHost, ctx->stream));
            MKID_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            MKID_REQUIRE(ctx, flag_h == 0, "decode_lists: more than 4 end-of-second words inside one range of the input");
            const DecRange *hr = (const DecRange *)ctx->dec_ranges_host.data();
            std::vector<int32_t> roach_first(cfg->n_roaches + 1, n_ranges);
            for (int k = n_ranges - 1; k >= 0; --k) roach_first[hr[k].roach] = k;
            for (int q = cfg->n_roaches - 1; q >= 0; --q) roach_first[q] = std::min(roach_first[q], roach_first[q + 1]);
            const size_t n_keys = (size_t)cfg->exptime * (lists->by_roach ? (size_t)cfg->n_roaches : (size_t)n_pix);
            char *lbuf;
            if ((rc = mkid_scratch(ctx, SCR_AUX5, n_keys * 4 + (size_t)(cfg->n_roaches + 1) * 4 + 64, (void **)&lbuf))) return rc;
            uint32_t *d_acc = (uint32_t *)lbuf;
            int32_t *d_rf = (int32_t *)(lbuf + n_keys * 4);
            MKID_CUDA(ctx, cudaMemsetAsync(d_acc, 0, n_keys * 4, ctx->stream));
            MKID_CUDA(ctx, cudaMemcpyAsync(d_rf, roach_first.data(), roach_first.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
            MKID_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            void *d_lw, *d_lo;
            if ((rc = mkid_stage_out(ctx, lists->list_words, (size_t)lists->list_cap * 8, SCR_OUT2, false, &d_lw))) return rc;
            if ((rc = mkid_stage_out(ctx, lists->list_offsets, (n_keys + 1) * 8, SCR_OUT3, false, &d_lo))) return rc;
            ListParams lp;
            lp.words = p.words; lp.wire = p.wire; lp.ranges = d_ranges; lp.rout = d_rout; lp.base = d_base; lp.eos_tot = d_eos; lp.roach_first = d_rf;
            lp.rows = d_rows; lp.acc = d_acc; lp.offsets = (long long *)d_lo; lp.out = (uint64_t *)d_lw; lp.out_cap = lists->list_cap;
            lp.n_ranges = n_ranges; lp.n_roaches = cfg->n_roaches; lp.n_pix = (int)n_pix; lp.npix_per_roach = cfg->npix_per_roach;
            lp.exptime = cfg->exptime; lp.cap = cfg->max_events - 1; lp.flag = d_flag;
            cudaEvent_t *lev = ctx->dbg_events + 4;
            auto lmark = [&](int i) { if (!timing) return; if (!lev[i]) cudaEventCreate(&lev[i]); cudaEventRecord(lev[i], ctx->stream); };
            uint32_t *d_rsum = nullptr;
            if (lists->by_roach) {
                if ((rc = mkid_scratch(ctx, SCR_AUX2, (size_t)n_ranges * DEC_MAX_LS * 4 + 64, (void **)&d_rsum))) return rc;
                lp.n_pix = cfg->n_roaches;              // the scan kernels run over exptime * n_pix keys
                lp.cap = 0x7FFFFFFF;
            }
            lmark(0);
            if (lists->by_roach) {
                merge_rangesum_kernel<<<(n_ranges + 7) / 8, 256, 0, ctx->stream>>>(lp, d_rsum);
                MKID_CHECK_LAUNCH(ctx);
                merge_rowstart_kernel<<<cfg->n_roaches, 256, 0, ctx->stream>>>(lp, d_rsum);
            } else {
                list_rowstart_kernel<<<cfg->n_roaches, 256, 0, ctx->stream>>>(lp);
            }
            MKID_CHECK_LAUNCH(ctx);
            lmark(1);
            const int n_sb = (int)((n_keys + LIST_SCAN_BLOCK - 1) / LIST_SCAN_BLOCK);
            long long *d_bs;
            if ((rc = mkid_scratch(ctx, SCR_AUX3, (size_t)n_sb * 8 + 64, (void **)&d_bs))) return rc;
            list_blocksum_kernel<<<n_sb, 1024, 0, ctx->stream>>>(lp, d_bs);
            MKID_CHECK_LAUNCH(ctx);
            list_blockscan_kernel<<<1, 1024, 0, ctx->stream>>>(lp, d_bs, n_sb);
            MKID_CHECK_LAUNCH(ctx);
            list_offsets_kernel<<<n_sb, 1024, 0, ctx->stream>>>(lp, d_bs);
            MKID_CHECK_LAUNCH(ctx);
            lmark(2);
            if (lists->by_roach) {
                const int mgrid = (n_ranges + LIST_WARPS - 1) / LIST_WARPS;
                if (wire_fmt) merge_scatter_kernel<true><<<mgrid, LIST_WARPS * 32, 0, ctx->stream>>>(lp, d_rsum);
                else merge_scatter_kernel<false><<<mgrid, LIST_WARPS * 32, 0, ctx->stream>>>(lp, d_rsum);
            } else {
                if (wire_fmt) list_scatter_kernel<true><<<n_ranges, LIST_WARPS * 32, 0, ctx->stream>>>(lp);
                else list_scatter_kernel<false><<<n_ranges, LIST_WARPS * 32, 0, ctx->stream>>>(lp);
            }
            MKID_CHECK_LAUNCH(ctx);
            lmark(3);
            if (timing) {
                float t[3];
                cudaEventSynchronize(lev[3]);
                for (int i = 0; i < 3; ++i) cudaEventElapsedTime(&t[i], lev[i], lev[i + 1]);
                fprintf(stderr, "[mkid decode timing] lists: rowstart %.1f us  offsets %.1f us  scatter %.1f us\n", t[0] * 1e3f, t[1] * 1e3f, t[2] * 1e3f);
            }
            MKID_CUDA(ctx, cudaMemcpyAsync(&flag_h, d_flag, 4, cudaMemcpyDeviceToHost, ctx->stream));
            MKID_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            MKID_REQUIRE(ctx, (flag_h & 2) == 0, "decode_lists: list_words capacity too small (see list_offsets[last])");
            if ((rc = mkid_stage_out_finish(ctx, lists->list_words, (size_t)lists->list_cap * 8, d_lw))) return rc;
            if ((rc = mkid_stage_out_finish(ctx, lists->list_offsets, (n_keys + 1) * 8, d_lo))) return rc;
        }
    }
    if (lists && n_ranges == 0) {          // no words: every list is empty
        void *d_lo;
        const size_t ob = ((size_t)cfg->exptime * (lists->by_roach ? (size_t)cfg->n_roaches : (size_t)n_pix) + 1) * 8;
        if ((rc = mkid_stage_out(ctx, lists->list_offsets, ob, SCR_OUT3, false, &d_lo))) return rc;
        MKID_CUDA(ctx, cudaMemsetAsync(d_lo, 0, ob, ctx->stream));
        if ((rc = mkid_stage_out_finish(ctx, lists->list_offsets, ob, d_lo))) return rc;
    }
    rc = mkid_stage_out_finish(ctx, counts_raw, counts_bytes, d_counts);
    if (rc) return rc;
    if (want_hist) {
        rc = mkid_stage_out_finish(ctx, hist, hist_bytes, d_hist);
        if (rc) return rc;
    }
    if (seg_sec_out) {
        std::vector<int32_t> tmp(n_seg, 0);
        if (n_ranges > 0) {
            MKID_CUDA(ctx, cudaMemcpyAsync(tmp.data(), d_sec_out, (size_t)n_seg * 4, cudaMemcpyDeviceToHost, ctx->stream));
            MKID_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        }
        for (int i = 0; i < n_seg; ++i) seg_sec_out[i] = seg_chunks[i] == 0 ? sec0[i] : tmp[i];
    }
    if (stats) {
        unsigned long long h[5];
        MKID_CUDA(ctx, cudaMemcpyAsync(h, d_stats, 40, cudaMemcpyDeviceToHost, ctx->stream));
        MKID_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (mkid_is_device_ptr(stats)) {
            mkid_decode_stats cur;
            MKID_CUDA(ctx, cudaMemcpy(&cur, stats, sizeof(cur), cudaMemcpyDeviceToHost));
            cur.n_eos += h[0]; cur.n_corrupt_eos += h[1]; cur.n_nonpixel += h[2]; cur.n_ignored += h[3]; cur.n_valid += h[4];
            MKID_CUDA(ctx, cudaMemcpy(stats, &cur, sizeof(cur), cudaMemcpyHostToDevice));
        } else {
            stats->n_eos += h[0]; stats->n_corrupt_eos += h[1]; stats->n_nonpixel += h[2];
            stats->n_ignored += h[3]; stats->n_valid += h[4];
        }
    }
    return MKID_OK;
}

}  // namespace

extern "C" int mkid_decode_words(mkid_ctx *ctx, const uint64_t *words, int64_t n_words, const int64_t *seg_offset,
                                 const int32_t *seg_roach, const int32_t *seg_sec, int32_t *seg_sec_out,
                                 int32_t n_segments, const mkid_decode_cfg *cfg, uint32_t *counts_raw,
                                 uint32_t *hist, mkid_decode_stats *stats) {
    if (!ctx) return MKID_EINVAL;
    MKID_REQUIRE(ctx, words || n_words == 0, "words is NULL");
    return decode_common(ctx, words, nullptr, n_words, seg_offset, nullptr, seg_roach, seg_sec, seg_sec_out, n_segments, cfg,
                         counts_raw, hist, stats);
}

extern "C" int mkid_decode_words_seg(mkid_ctx *ctx, const uint64_t *words, int64_t n_words, const int64_t *seg_start,
                                     const int64_t *seg_len, const int32_t *seg_roach, const int32_t *seg_sec,
                                     int32_t *seg_sec_out, int32_t n_segments, const mkid_decode_cfg *cfg,
                                     uint32_t *counts_raw, uint32_t *hist, mkid_decode_stats *stats) {
    if (!ctx) return MKID_EINVAL;
    MKID_REQUIRE(ctx, (words || n_words == 0) && seg_len, "words / seg_len is NULL");
    return decode_common(ctx, words, nullptr, n_words, seg_start, seg_len, seg_roach, seg_sec, seg_sec_out, n_segments, cfg,
                         counts_raw, hist, stats);
}

extern "C" int mkid_decode_words_dev(mkid_ctx *ctx, const uint64_t *words, int64_t n_words, const int64_t *seg_start,
                                     const int64_t *seg_cap, const int32_t *seg_len_dev, const int32_t *seg_roach,
                                     const int32_t *seg_sec_dev, int32_t *seg_sec_out_dev, int32_t n_segments,
                                     const mkid_decode_cfg *cfg, uint32_t *counts_raw, uint32_t *hist) {
    if (!ctx) return MKID_EINVAL;
    MKID_REQUIRE(ctx, words && seg_cap && seg_len_dev && seg_sec_dev && seg_sec_out_dev, "decode_words_dev: NULL argument");
    MKID_REQUIRE(ctx, mkid_is_device_ptr(words) && mkid_is_device_ptr(seg_len_dev) && mkid_is_device_ptr(seg_sec_dev) &&
                          mkid_is_device_ptr(seg_sec_out_dev) && seg_sec_dev != seg_sec_out_dev &&
                          mkid_is_device_ptr(counts_raw) && (!hist || mkid_is_device_ptr(hist)),
                 "decode_words_dev: words, lengths, second counters (distinct in/out) and products must be device memory");
    for (int i = 0; i < n_segments; ++i) MKID_REQUIRE(ctx, seg_cap[i] > 0, "decode_words_dev: empty segment capacity");
    return decode_common(ctx, words, nullptr, n_words, seg_start, seg_cap, seg_roach, nullptr, nullptr, n_segments, cfg,
                         counts_raw, hist, nullptr, seg_len_dev, seg_sec_dev, seg_sec_out_dev);
}

extern "C" int mkid_merge_words_dev(mkid_ctx *ctx, const uint64_t *words, const int64_t *seg_start, const int64_t *seg_cap,
                                    const int32_t *seg_len_dev, const int32_t *seg_sec_dev, int32_t n_segments,
                                    const mkid_decode_cfg *cfg, uint64_t *list_words, int64_t list_cap, int32_t *list_offsets) {
    if (!ctx) return MKID_EINVAL;
    MKID_REQUIRE(ctx, words && seg_start && seg_cap && seg_len_dev && seg_sec_dev && cfg && list_words && list_offsets && list_cap > 0,
                 "merge_words_dev: NULL argument");
    MKID_REQUIRE(ctx, n_segments >= 1 && n_segments <= 1024, "merge_words_dev: 1..1024 segments");
    MKID_REQUIRE(ctx, mkid_is_device_ptr(words) && mkid_is_device_ptr(seg_len_dev) && mkid_is_device_ptr(seg_sec_dev) &&
                          mkid_is_device_ptr(list_words) && mkid_is_device_ptr(list_offsets),
                 "merge_words_dev: words, lengths, second counters and outputs must be device memory");
    MKID_CUDA(ctx, cudaSetDevice(ctx->device));
    // segment table + per-key counts in a private scratch slot (re-uploaded only when it changes)
    std::vector<DecSeg> segs(n_segments);
    for (int i = 0; i < n_segments; ++i) segs[i] = DecSeg{(long long)seg_start[i], (int)std::min<int64_t>(seg_cap[i], INT32_MAX), i};
    const size_t tab_bytes = segs.size() * sizeof(DecSeg), cnt_bytes = (size_t)MERGE_MAX_SEC * n_segments * 4;
    const size_t need = ((tab_bytes + 15) / 16) * 16 + cnt_bytes;
    if (ctx->merge_bytes < need) {
        if (ctx->merge_dev) { MKID_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); MKID_CUDA(ctx, cudaFree(ctx->merge_dev)); ctx->merge_dev = nullptr; }
        if (cudaMalloc(&ctx->merge_dev, need) != cudaSuccess) return mkid_fail(ctx, MKID_ENOMEM, "merge_words_dev: device allocation failed");
        ctx->merge_bytes = need;
        ctx->merge_host.clear();
    }
    void *scr = ctx->merge_dev;
    std::vector<char> key((const char *)segs.data(), (const char *)segs.data() + tab_bytes);
    if (key != ctx->merge_host) {                          // (re-uploaded only when the table changes: no host round trip per batch)
        MKID_CUDA(ctx, cudaMemcpyAsync(scr, segs.data(), tab_bytes, cudaMemcpyHostToDevice, ctx->stream));
        MKID_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        ctx->merge_host = key;
    }
    MergeParams p;
    p.words = words; p.seg_len_dev = seg_len_dev; p.seg_sec_dev = seg_sec_dev; p.n_seg = n_segments;
    p.npix_per_roach = cfg->npix_per_roach; p.exptime = cfg->exptime;
    p.cnt = (int32_t *)((char *)scr + ((tab_bytes + 15) / 16) * 16);
    p.out = list_words; p.out_cap = list_cap; p.offsets = list_offsets;
    merge_small_kernel<false><<<n_segments, 1024, 0, ctx->stream>>>(p, (const DecSeg *)scr);
    MKID_CHECK_LAUNCH(ctx);
    merge_small_kernel<true><<<n_segments, 1024, 0, ctx->stream>>>(p, (const DecSeg *)scr);
    MKID_CHECK_LAUNCH(ctx);
    return MKID_OK;
}

extern "C" int mkid_decode_lists(mkid_ctx *ctx, const uint64_t *words, int64_t n_words, const int64_t *seg_offset,
                                 const int32_t *seg_roach, const int32_t *seg_sec, int32_t *seg_sec_out, int32_t n_segments,
                                 const mkid_decode_cfg *cfg, uint32_t *counts_raw, uint64_t *list_words, int64_t list_cap,
                                 int64_t *list_offsets, mkid_decode_stats *stats) {
    if (!ctx) return MKID_EINVAL;
    MKID_REQUIRE(ctx, (words || n_words == 0) && list_words && list_offsets && list_cap > 0 && cfg && cfg->max_events >= 2,
                 "decode_lists: NULL argument");
    ListRequest lr{list_words, list_cap, list_offsets, false};
    return decode_common(ctx, words, nullptr, n_words, seg_offset, nullptr, seg_roach, seg_sec, seg_sec_out, n_segments, cfg,
                         counts_raw, nullptr, stats, nullptr, nullptr, nullptr, &lr);
}

extern "C" int mkid_decode_merged(mkid_ctx *ctx, const uint64_t *words, int64_t n_words, const int64_t *seg_offset,
                                  const int32_t *seg_roach, const int32_t *seg_sec, int32_t *seg_sec_out, int32_t n_segments,
                                  const mkid_decode_cfg *cfg, uint32_t *counts_raw, uint64_t *list_words, int64_t list_cap,
                                  int64_t *list_offsets, mkid_decode_stats *stats) {
    if (!ctx) return MKID_EINVAL;
    MKID_REQUIRE(ctx, (words || n_words == 0) && list_words && list_offsets && list_cap > 0 && cfg, "decode_merged: NULL argument");
    ListRequest lr{list_words, list_cap, list_offsets, true};
    return decode_common(ctx, words, nullptr, n_words, seg_offset, nullptr, seg_roach, seg_sec, seg_sec_out, n_segments, cfg,
                         counts_raw, nullptr, stats, nullptr, nullptr, nullptr, &lr);
}

extern "C" int mkid_decode_wire_lists(mkid_ctx *ctx, const uint32_t *wire, int64_t n_bundles, const int64_t *seg_offset,
                                      const int32_t *seg_roach, const int32_t *seg_sec, int32_t *seg_sec_out, int32_t n_segments,
                                      const mkid_decode_cfg *cfg, uint32_t *counts_raw, int32_t merged, uint64_t *list_words,
                                      int64_t list_cap, int64_t *list_offsets, mkid_decode_stats *stats) {
    if (!ctx) return MKID_EINVAL;
    MKID_REQUIRE(ctx, wire && list_words && list_offsets && list_cap > 0 && cfg && (merged || cfg->max_events >= 2),
                 "decode_wire_lists: NULL argument");
    ListRequest lr{list_words, list_cap, list_offsets, merged != 0};
    return decode_common(ctx, nullptr, wire, n_bundles, seg_offset, nullptr, seg_roach, seg_sec, seg_sec_out, n_segments, cfg,
                         counts_raw, nullptr, stats, nullptr, nullptr, nullptr, &lr);
}

extern "C" int mkid_decode_wire(mkid_ctx *ctx, const uint32_t *wire, int64_t n_bundles, const int64_t *seg_offset,
                                const int32_t *seg_roach, const int32_t *seg_sec, int32_t *seg_sec_out,
                                int32_t n_segments, const mkid_decode_cfg *cfg, uint32_t *counts_raw, uint32_t *hist,
                                mkid_decode_stats *stats) {
    if (!ctx) return MKID_EINVAL;
    MKID_REQUIRE(ctx, wire, "wire is NULL");
    return decode_common(ctx, nullptr, wire, n_bundles, seg_offset, nullptr, seg_roach, seg_sec, seg_sec_out, n_segments, cfg,
                         counts_raw, hist, stats);
}

extern "C" int mkid_counts_cap(mkid_ctx *ctx, const uint32_t *counts_raw, uint32_t *counts, int64_t n,
                               int32_t max_events) {
    if (!ctx) return MKID_EINVAL;
    MKID_REQUIRE(ctx, counts_raw && counts && n >= 0 && max_events >= 1, "bad counts_cap args");
    if (n == 0) return MKID_OK;
    MKID_CUDA(ctx, cudaSetDevice(ctx->device));
    const void *d_in; void *d_out;
    int rc = mkid_stage_in(ctx, counts_raw, n * 4, SCR_IN, &d_in); if (rc) return rc;
    if (counts == counts_raw && !mkid_is_device_ptr(counts)) d_out = (void *)d_in;
    else { rc = mkid_stage_out(ctx, counts, n * 4, SCR_OUT0, false, &d_out); if (rc) return rc; }
    int grid = (int)std::min<int64_t>((n + 255) / 256, (int64_t)ctx->num_sms * 8);
    counts_cap_kernel<<<grid, 256, 0, ctx->stream>>>((const uint32_t *)d_in, (uint32_t *)d_out, n, (uint32_t)(max_events - 1));
    MKID_CHECK_LAUNCH(ctx);
    return mkid_stage_out_finish(ctx, counts, n * 4, d_out);
}

extern "C" int mkid_unpack_fields(mkid_ctx *ctx, const uint64_t *words, int64_t n, uint8_t *ch, uint32_t *ts,
                                  uint16_t *base, uint16_t *peak, uint16_t *p1) {
    if (!ctx) return MKID_EINVAL;
    if (n == 0) return MKID_OK;
    MKID_REQUIRE(ctx, words && n > 0, "bad unpack args");
    MKID_CUDA(ctx, cudaSetDevice(ctx->device));
    const void *d_in; void *d_ch, *d_ts, *d_base, *d_peak, *d_p1;
    int rc;
    if ((rc = mkid_stage_in(ctx, words, n * 8, SCR_IN, &d_in))) return rc;
    if ((rc = mkid_stage_out(ctx, ch, n, SCR_OUT0, false, &d_ch))) return rc;
    if ((rc = mkid_stage_out(ctx, ts, n * 4, SCR_OUT1, false, &d_ts))) return rc;
    if ((rc = mkid_stage_out(ctx, base, n * 2, SCR_OUT2, false, &d_base))) return rc;
    if ((rc = mkid_stage_out(ctx, peak, n * 2, SCR_OUT3, false, &d_peak))) return rc;
    if ((rc = mkid_stage_out(ctx, p1, n * 2, SCR_AUX0, false, &d_p1))) return rc;
    int grid = (int)std::min<int64_t>((n + 255) / 256, (int64_t)ctx->num_sms * 16);
    unpack_fields_kernel<<<grid, 256, 0, ctx->stream>>>((const uint64_t *)d_in, n, (uint8_t *)d_ch, (uint32_t *)d_ts,
                                                        (uint16_t *)d_base, (uint16_t *)d_peak, (uint16_t *)d_p1);
    MKID_CHECK_LAUNCH(ctx);
    if ((rc = mkid_stage_out_finish(ctx, ch, n, d_ch))) return rc;
    if ((rc = mkid_stage_out_finish(ctx, ts, n * 4, d_ts))) return rc;
    if ((rc = mkid_stage_out_finish(ctx, base, n * 2, d_base))) return rc;
    if ((rc = mkid_stage_out_finish(ctx, peak, n * 2, d_peak))) return rc;
    return mkid_stage_out_finish(ctx, p1, n * 2, d_p1);
}

extern "C" int mkid_reinterpret_bin(mkid_ctx *ctx, const uint64_t *values, int64_t n, int32_t n_bits,
                                    int32_t binary_point, int32_t n_bits_after_end, double *out) {
    if (!ctx) return MKID_EINVAL;
    if (n == 0) return MKID_OK;
    MKID_REQUIRE(ctx, values && out && n > 0 && n_bits >= 1 && n_bits <= 63 && n_bits_after_end >= 0 &&
                          n_bits_after_end < 64, "bad reinterpret_bin args");
    MKID_CUDA(ctx, cudaSetDevice(ctx->device));
    const void *d_in; void *d_out; int rc;
    if ((rc = mkid_stage_in(ctx, values, n * 8, SCR_IN, &d_in))) return rc;
    if ((rc = mkid_stage_out(ctx, out, n * 8, SCR_OUT0, false, &d_out))) return rc;
    int grid = (int)std::min<int64_t>((n + 255) / 256, (int64_t)ctx->num_sms * 16);
    reinterpret_bin_kernel<<<grid, 256, 0, ctx->stream>>>((const uint64_t *)d_in, n, n_bits, binary_point,
                                                          n_bits_after_end, (double *)d_out);
    MKID_CHECK_LAUNCH(ctx);
    return mkid_stage_out_finish(ctx, out, n * 8, d_out);
}

extern "C" int mkid_quicklook_image(mkid_ctx *ctx, const uint32_t *counts_sec, const int32_t *pixel_adr,
                                    int32_t n, uint16_t *image) {
    if (!ctx) return MKID_EINVAL;
    MKID_REQUIRE(ctx, counts_sec && pixel_adr && image && n > 0, "bad quicklook args");
    MKID_REQUIRE(ctx, mkid_is_device_ptr(counts_sec), "counts_sec must be device memory (its length is not passed)");
    MKID_CUDA(ctx, cudaSetDevice(ctx->device));
    const void *d_adr; void *d_img; int rc;
    if ((rc = mkid_stage_in(ctx, pixel_adr, (size_t)n * 4, SCR_IN, &d_adr))) return rc;
    if ((rc = mkid_stage_out(ctx, image, (size_t)n * 2, SCR_OUT0, false, &d_img))) return rc;
    quicklook_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(counts_sec, (const int32_t *)d_adr, n, (uint16_t *)d_img);
    MKID_CHECK_LAUNCH(ctx);
    return mkid_stage_out_finish(ctx, image, (size_t)n * 2, d_img);
}
