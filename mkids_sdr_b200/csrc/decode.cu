// K6: photon-word decode + per-(second,pixel) binning + per-pixel pulse-height histogram.
//
// Replaces the receive/bin loop of DataReadout/ReadoutControls/lib/PacketMaster.c:286-397 and
// the per-word unpack of DataReadout/ChannelizerControls/ROACH_Pulses.py:795-832.
//
// HBM-bound integer work: every word is read exactly once (8 B/word).  A chunk is 4096 words
// (32 KiB = half a PulseServer bundle, PacketMaster.c:42-44).  One persistent, warp-specialised CTA
// per SM:
//   6 scout warps        one per stage of a 6 x 32 KiB shared-memory ring: take a chunk from an ordered
//                        ticket, stream it in with 1-D TMA bulk copies (mbarrier completion), count its
//                        end-of-second words as soon as it has landed, publish that count and resolve
//                        "seconds closed before this chunk" with a decoupled look-back over the published
//                        counts -- each scout has a full ring revolution to hide its L2 round trips;
//   16 worker warps      unpack the bitfields from shared memory and bin: per-(second,pixel) counts
//                        through a per-chunk shared-memory histogram (double buffered, one named
//                        barrier per chunk), pulse-height histogram through shared memory when it is
//                        small, else global reductions.
#include <stdlib.h>

#include "common.cuh"

namespace {

constexpr int DEC_WORKERS = 512;                   // worker threads
constexpr int DEC_THREADS = DEC_WORKERS + 32 * 6;  // + one producer/scout warp per ring stage
constexpr int DEC_CHUNK = 4096;                    // words per chunk (wire format: half a bundle)
constexpr int DEC_BUNDLE = 8192;                   // PacketMaster.c:44 BUFSIZE_INTS
constexpr int DEC_WPT = DEC_CHUNK / DEC_WORKERS;   // 8 words per worker thread
constexpr int DEC_STAGES = 6;
constexpr int DEC_STAGE_BYTES = DEC_CHUNK * 8;
constexpr int DEC_SMEM_HIST = 4096;                // smem-privatised histogram entries
constexpr int DEC_MAX_EOS = 64;                    // end-of-second positions kept per chunk (more: slow path)

struct DecParams {
    const uint64_t *words;     // flat format (or nullptr)
    const uint32_t *wire;      // wire format (or nullptr)
    const int64_t *seg_first_chunk;   // [n_seg+1]
    const int64_t *seg_offset;        // [n_seg] first word (flat) / bundle (wire) of each segment
    const int64_t *seg_len;           // [n_seg] words / bundles
    const int32_t *seg_roach;
    const int32_t *seg_sec;
    int32_t *seg_sec_out;
    int32_t n_seg;
    int64_t n_chunks;
    int32_t n_roaches, npix_per_roach, exptime;
    int32_t field_shift, n_bins;
    const uint16_t *bin_lut;
    uint32_t *counts;          // [exptime][n_pix]
    uint32_t *hist;            // [n_pix][n_bins]
    unsigned long long *stats; // 5 x u64
    unsigned long long *state; // [n_chunks] look-back records
    unsigned int *ticket;
    unsigned long long *prof;  // optional [8] cycle accumulators (MKID_DEC_PROFILE=1): see decode_common
};

struct StageMeta {
    long long chunk;       // < 0: no more work
    int roach, base_sec, n_here, n_eos, last_chunk, seg;
    int eos_pos[DEC_MAX_EOS];
};

__device__ __forceinline__ uint32_t bswap32(uint32_t x) { return __byte_perm(x, 0, 0x0123); }
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(mk_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void workers_sync() { asm volatile("bar.sync 1, %0;" ::"n"(DEC_WORKERS) : "memory"); }

template <bool WIRE, bool SMEM_HIST>
__global__ void __launch_bounds__(DEC_THREADS, 1) decode_kernel(DecParams p) {
    extern __shared__ __align__(128) unsigned char s_ring[];      // DEC_STAGES x 64 KiB
    __shared__ uint32_t s_cnt[2][256];
    __shared__ uint32_t s_hist[SMEM_HIST ? DEC_SMEM_HIST : 1];
    __shared__ uint16_t s_lut[4096];
    __shared__ unsigned long long s_stat[5];
    __shared__ __align__(8) uint64_t s_full[DEC_STAGES], s_ready[DEC_STAGES], s_empty[DEC_STAGES];
    __shared__ StageMeta s_meta[DEC_STAGES];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool use_lut = p.bin_lut != nullptr;
    if (use_lut)
        for (int i = tid; i < 4096; i += DEC_THREADS) s_lut[i] = p.bin_lut[i];
    if (tid < 5) s_stat[tid] = 0;
    for (int i = tid; i < 512; i += DEC_THREADS) (&s_cnt[0][0])[i] = 0;
    if (SMEM_HIST)
        for (int i = tid; i < DEC_SMEM_HIST; i += DEC_THREADS) s_hist[i] = 0;
    if (tid == 0) {
        for (int i = 0; i < DEC_STAGES; ++i) {
            mk_mbar_init(&s_full[i], 1);
            mk_mbar_init(&s_ready[i], 1);
            mk_mbar_init(&s_empty[i], DEC_WORKERS / 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int n_pix = p.n_roaches * p.npix_per_roach;

    if (warp >= DEC_WORKERS / 32) {
        // =========================== producer / scout warps ===========================
        // scout j owns ring stage j: wait until the workers freed it, take the next ticket, start the bulk
        // load, scan the chunk when it has landed, resolve its second index, hand it to the workers.
        const int ss = warp - DEC_WORKERS / 32;
        StageMeta &m = s_meta[ss];
        for (int r = 0;; ++r) {                              // r-th use of this stage
            // (the ticket must not be drawn before the stage is free: a held ticket stalls the look-back of
            // every later chunk)
            long long t0 = p.prof ? clock64() : 0;
            if (r > 0) mk_mbar_wait(&s_empty[ss], (r - 1) & 1);
            long long t1 = p.prof ? clock64() : 0;
            long long c = 0;
            if (lane == 0) c = (long long)atomicAdd(p.ticket, 1u);
            c = __shfl_sync(0xffffffffu, c, 0);
            long long t2 = p.prof ? clock64() : 0;
            if (c >= p.n_chunks) {                           // no more work for this stage, ever (sticky)
                if (lane == 0) { m.chunk = -1; mbar_arrive(&s_ready[ss]); }
                break;
            }
            {
                int lo = 0, hi = p.n_seg;       // segment of this chunk: last g with first_chunk[g] <= c
                while (hi - lo > 1) {
                    const int mid = (lo + hi) >> 1;
                    if (p.seg_first_chunk[mid] <= c) lo = mid; else hi = mid;
                }
                const int g = lo;
                const long long lc = c - p.seg_first_chunk[g];
                const unsigned char *src;
                int n_here = DEC_CHUNK;
                if (WIRE) {   // chunk lc = half (lc & 1) of bundle lc >> 1: its low halves; the high halves are 32 KiB further
                    src = reinterpret_cast<const unsigned char *>(p.wire) + (size_t)(p.seg_offset[g] + (lc >> 1)) * (DEC_BUNDLE * 8) +
                          (size_t)(lc & 1) * (DEC_CHUNK * 4);
                } else {
                    const long long rem = p.seg_len[g] - lc * DEC_CHUNK;
                    n_here = rem < DEC_CHUNK ? (int)rem : DEC_CHUNK;
                    src = reinterpret_cast<const unsigned char *>(p.words + p.seg_offset[g] + lc * DEC_CHUNK);
                }
                unsigned char *dst = s_ring + (size_t)ss * DEC_STAGE_BYTES;
                const bool tma = n_here == DEC_CHUNK && (reinterpret_cast<uintptr_t>(src) & 15) == 0;
                if (lane == 0) {
                    m.chunk = c; m.seg = g; m.roach = p.seg_roach[g]; m.n_here = n_here;
                    m.last_chunk = (c + 1 == p.seg_first_chunk[g + 1]);
                    if (tma) {
                        mk_mbar_expect_tx(&s_full[ss], DEC_STAGE_BYTES);
                        if (WIRE) {
                            mk_bulk_g2s(dst, src, DEC_STAGE_BYTES / 2, &s_full[ss]);
                            mk_bulk_g2s(dst + DEC_STAGE_BYTES / 2, src + DEC_BUNDLE * 4, DEC_STAGE_BYTES / 2, &s_full[ss]);
                        } else {
#pragma unroll
                            for (int q = 0; q < 2; ++q)
                                mk_bulk_g2s(dst + q * (DEC_STAGE_BYTES / 2), src + q * (DEC_STAGE_BYTES / 2),
                                            DEC_STAGE_BYTES / 2, &s_full[ss]);
                        }
                    }
                }
                if (!tma) {   // ragged tail / odd alignment (flat format only): the warp copies it, zero padded
                    const uint64_t *sw = reinterpret_cast<const uint64_t *>(src);
                    uint64_t *dw = reinterpret_cast<uint64_t *>(dst);
                    for (int i = lane; i < DEC_CHUNK; i += 32) dw[i] = i < n_here ? sw[i] : 0ull;
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&s_full[ss]);
                }
            }
            __syncwarp();
            mk_mbar_wait(&s_full[ss], r & 1);
            long long t3 = p.prof ? clock64() : 0;
            // count end-of-second words (channel field 255) and remember their positions, in order.
            // Fast path: 16 x 128-bit loads per lane in flight, one ballot per batch; the ordered position
            // list is only built for batches that contain an end-of-second word (rare).
            const unsigned char *base = s_ring + (size_t)ss * DEC_STAGE_BYTES;
            const uint4 *scan = reinterpret_cast<const uint4 *>(WIRE ? base + DEC_STAGE_BYTES / 2 : base);
            constexpr int SCAN_IT = (WIRE ? DEC_STAGE_BYTES / 2 : DEC_STAGE_BYTES) / 16 / 32;   // uint4 iterations per lane
            constexpr int POS_PER_IT = WIRE ? 128 : 64;                                       // positions covered by one iteration
            int n_eos = 0;
            constexpr int SB = 16;                           // 128-bit loads in flight per lane
            for (int i0 = 0; i0 < SCAN_IT; i0 += SB) {
                uint4 v[SB];
#pragma unroll
                for (int u = 0; u < SB; ++u) v[u] = scan[(i0 + u) * 32 + lane];
                bool any = false;
#pragma unroll
                for (int u = 0; u < SB; ++u) {
                    if (WIRE) any |= ((v[u].x & 0xFFu) == 0xFFu) | ((v[u].y & 0xFFu) == 0xFFu) | ((v[u].z & 0xFFu) == 0xFFu) | ((v[u].w & 0xFFu) == 0xFFu);
                    else any |= ((v[u].y >> 24) == 0xFFu) | ((v[u].w >> 24) == 0xFFu);
                }
                if (__ballot_sync(0xffffffffu, any)) {
                    // position-ordered pass over this batch (zero padding beyond n_here is never an EOS word)
                    const int p0 = i0 * POS_PER_IT, p1 = p0 + SB * POS_PER_IT;
                    for (int pb = p0; pb < p1; pb += 32) {
                        const int pos = pb + lane;
                        bool is_eos;
                        if (WIRE) is_eos = (reinterpret_cast<const uint32_t *>(base + DEC_STAGE_BYTES / 2)[pos] & 0xFFu) == 0xFFu;
                        else is_eos = (reinterpret_cast<const uint32_t *>(base)[2 * pos + 1] >> 24) == 0xFFu;
                        const unsigned bal = __ballot_sync(0xffffffffu, is_eos);
                        if (bal) {
                            const int at = n_eos + __popc(bal & ((1u << lane) - 1));
                            if (is_eos && at < DEC_MAX_EOS) m.eos_pos[at] = pos;
                            n_eos += __popc(bal);
                        }
                    }
                }
            }
            long long t4 = p.prof ? clock64() : 0;
            // decoupled look-back: seconds closed before this chunk
            const int g = m.seg;
            const long long lc = c - p.seg_first_chunk[g];
            int base_sec = 0;
            if (lc == 0) {
                base_sec = p.seg_sec[g];
            } else {
                // aggregate first (plain 64-bit store: self-contained record, nothing to wait for)
                if (lane == 0) *reinterpret_cast<volatile unsigned long long *>(&p.state[c]) = (1ull << 32) | (unsigned)n_eos;
                const long long first = c - lc;
                long long q = c - 1;
                int acc = 0;
                for (;;) {
                    // 128 predecessors per round: lane l looks at q - l - 32*j, j = 0..3 (nearest first)
                    unsigned long long sv[4];
                    bool valid[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const long long idx = q - lane - 32 * j;
                        valid[j] = idx >= first;
                        sv[j] = valid[j] ? *reinterpret_cast<volatile unsigned long long *>(&p.state[idx]) : 0ull;
                    }
                    bool stop_found = false;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        if (stop_found) break;
                        const long long idx = q - lane - 32 * j;
                        while (valid[j] && (sv[j] >> 32) == 0) sv[j] = *reinterpret_cast<volatile unsigned long long *>(&p.state[idx]);
                        const unsigned incl = __ballot_sync(0xffffffffu, valid[j] && (sv[j] >> 32) == 2);
                        int v;
                        if (incl) {
                            const int stop = __ffs(incl) - 1;           // nearest predecessor holding a prefix
                            v = (lane <= stop) ? (int)(unsigned)sv[j] : 0;
                            stop_found = true;
                        } else {
                            v = valid[j] ? (int)(unsigned)sv[j] : 0;
                        }
#pragma unroll
                        for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
                        acc += v;
                    }
                    if (stop_found) break;
                    q -= 128;
                }
                base_sec = acc;
            }
            if (lane == 0) {
                *reinterpret_cast<volatile unsigned long long *>(&p.state[c]) = (2ull << 32) | (unsigned)(base_sec + n_eos);
                m.base_sec = base_sec;
                m.n_eos = n_eos;
                if (m.last_chunk && p.seg_sec_out) p.seg_sec_out[g] = base_sec + n_eos;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_ready[ss]);       // release: meta + data are ready for the workers
            if (p.prof && lane == 0) {
                const long long t5 = clock64();
                atomicAdd(&p.prof[0], (unsigned long long)(t1 - t0));   // scout: wait for a free stage
                atomicAdd(&p.prof[1], (unsigned long long)(t2 - t1));   // ticket
                atomicAdd(&p.prof[2], (unsigned long long)(t3 - t2));   // load issue + data arrival
                atomicAdd(&p.prof[3], (unsigned long long)(t4 - t3));   // EOS scan
                atomicAdd(&p.prof[4], (unsigned long long)(t5 - t4));   // look-back + publish
                atomicAdd(&p.prof[7], 1ull);
            }
        }
    } else {
        // =========================== worker warps ===========================
        unsigned n_eos_t = 0, n_bad = 0, n_nonpix = 0, n_ign = 0, n_ok = 0;
        // stages are visited round-robin; a stage whose scout ran out of tickets is dead for good.  The
        // scouts may draw their first tickets in any order, so a dead stage does not end the loop: it
        // ends when all stages are dead.
        int dead = 0, n_proc = 0;
        for (int k = 0; dead != (1 << DEC_STAGES) - 1; ++k) {
            const int st = k % DEC_STAGES;
            if ((dead >> st) & 1) continue;
            long long w0 = p.prof ? clock64() : 0;
            mk_mbar_wait(&s_ready[st], (k / DEC_STAGES) & 1);
            long long w1 = p.prof ? clock64() : 0;
            const StageMeta &m = s_meta[st];
            if (m.chunk < 0) { dead |= 1 << st; continue; }
            const int roach = m.roach, sec_base = m.base_sec, n_here = m.n_here, n_eos = m.n_eos;
            uint32_t *cnt = s_cnt[n_proc & 1];
            ++n_proc;
            const uint4 *sm = reinterpret_cast<const uint4 *>(s_ring + (size_t)st * DEC_STAGE_BYTES);
            if (n_here == DEC_CHUNK && n_eos == 0 && sec_base < p.exptime) {
                // fast path (almost every chunk): full chunk, no second boundary inside, all words live
                const bool f_hi = p.field_shift >= 32;
                const int f_sh = p.field_shift & 31;
                const int npix = p.npix_per_roach;
                uint32_t *hist_r = p.hist ? p.hist + (size_t)roach * npix * p.n_bins : nullptr;
#pragma unroll
                for (int i = 0; i < (WIRE ? DEC_WPT / 4 : DEC_WPT / 2); ++i) {
                    uint32_t whi[WIRE ? 4 : 2], wlo[WIRE ? 4 : 2];
                    if (WIRE) {
                        const uint4 l = sm[i * DEC_WORKERS + tid], h = sm[DEC_CHUNK / 4 + i * DEC_WORKERS + tid];
                        whi[0] = bswap32(h.x); whi[1] = bswap32(h.y); whi[2] = bswap32(h.z); whi[3] = bswap32(h.w);
                        wlo[0] = bswap32(l.x); wlo[1] = bswap32(l.y); wlo[2] = bswap32(l.z); wlo[3] = bswap32(l.w);
                    } else {
                        const uint4 v = sm[i * DEC_WORKERS + tid];
                        wlo[0] = v.x; whi[0] = v.y; wlo[1] = v.z; whi[1] = v.w;
                    }
#pragma unroll
                    for (int j = 0; j < (WIRE ? 4 : 2); ++j) {
                        const uint32_t hi = whi[j], lo = wlo[j];
                        const uint32_t adr = hi >> 24;
                        if ((int)adr >= npix) { ++n_nonpix; --n_ok; continue; }   // (adr == 255 cannot occur: n_eos == 0)
                        if (SMEM_HIST) {
                            // one shared atomic per word: row adr, column = bin or the overflow column n_bins;
                            // the per-pixel count is the row sum (taken at the flush)
                            const uint32_t f = ((f_hi ? hi : lo) >> f_sh) & 0xFFFu;
                            const uint32_t b = use_lut ? s_lut[f] : f;
                            atomicAdd(&s_hist[adr * (p.n_bins + 1) + min((int)b, p.n_bins)], 1u);
                        } else {
                            atomicAdd(&cnt[adr], 1u);
                            if (hist_r) {
                                const uint32_t f = ((f_hi ? hi : lo) >> f_sh) & 0xFFFu;
                                const uint32_t b = use_lut ? s_lut[f] : f;
                                if ((int)b < p.n_bins) atomicAdd(&hist_r[adr * p.n_bins + b], 1u);
                            }
                        }
                    }
                }
                n_ok += DEC_WPT;          // corrected by the non-pixel words below
            } else {
                // 16 words per thread, 16 B per lane per access; everything on 32-bit halves
                const bool f_hi = p.field_shift >= 32;
                const int f_sh = p.field_shift & 31;
    #pragma unroll
                for (int i = 0; i < (WIRE ? DEC_WPT / 4 : DEC_WPT / 2); ++i) {
                    uint32_t whi[WIRE ? 4 : 2], wlo[WIRE ? 4 : 2];
                    int pos0;
                    if (WIRE) {
                        const uint4 l = sm[i * DEC_WORKERS + tid], h = sm[DEC_CHUNK / 4 + i * DEC_WORKERS + tid];
                        whi[0] = bswap32(h.x); whi[1] = bswap32(h.y); whi[2] = bswap32(h.z); whi[3] = bswap32(h.w);
                        wlo[0] = bswap32(l.x); wlo[1] = bswap32(l.y); wlo[2] = bswap32(l.z); wlo[3] = bswap32(l.w);
                        pos0 = (i * DEC_WORKERS + tid) * 4;
                    } else {
                        const uint4 v = sm[i * DEC_WORKERS + tid];
                        wlo[0] = v.x; whi[0] = v.y; wlo[1] = v.z; whi[1] = v.w;
                        pos0 = (i * DEC_WORKERS + tid) * 2;
                    }
    #pragma unroll
                    for (int j = 0; j < (WIRE ? 4 : 2); ++j) {
                        const int pos = pos0 + j;
                        if (pos >= n_here) continue;
                        const uint32_t hi = whi[j], lo = wlo[j];
                        int l = 0;                                   // seconds closed inside the chunk before this word
                        if (n_eos) {
                            if (n_eos <= DEC_MAX_EOS) {
                                for (int e = 0; e < n_eos; ++e) l += m.eos_pos[e] < pos;
                            } else {                                 // pathological: rescan the chunk prefix
                                for (int q = 0; q < pos; ++q) {
                                    if (WIRE) l += (reinterpret_cast<const uint32_t *>(sm)[DEC_CHUNK + q] & 0xFFu) == 0xFFu;
                                    else l += (reinterpret_cast<const uint32_t *>(sm)[2 * q + 1] >> 24) == 0xFFu;
                                }
                            }
                        }
                        const int sec = sec_base + l;
                        const uint32_t adr = hi >> 24;
                        if (sec >= p.exptime) { ++n_ign; continue; }
                        if (adr == 255u) { ++n_eos_t; if ((hi & lo) != 0xFFFFFFFFu) ++n_bad; continue; }
                        if ((int)adr >= p.npix_per_roach) { ++n_nonpix; continue; }
                        ++n_ok;
                        uint32_t b = 0;
                        if (p.hist) {
                            const uint32_t f = ((f_hi ? hi : lo) >> f_sh) & 0xFFFu;
                            b = use_lut ? s_lut[f] : f;
                        }
                        if (SMEM_HIST && l == 0) {
                            atomicAdd(&s_hist[adr * (p.n_bins + 1) + min((int)b, p.n_bins)], 1u);
                        } else {
                            if (l == 0) atomicAdd(&cnt[adr], 1u);
                            else atomicAdd(&p.counts[(size_t)sec * n_pix + roach * p.npix_per_roach + adr], 1u);
                            if (p.hist && (int)b < p.n_bins)
                                atomicAdd(&p.hist[((size_t)(roach * p.npix_per_roach + adr)) * p.n_bins + b], 1u);
                        }
                    }
                }
            }
            // this thread no longer needs the ring slot
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_empty[st]);
            workers_sync();                                   // all counts of this chunk are in cnt[]
            if (sec_base < p.exptime && tid < p.npix_per_roach) {
                const uint32_t v = cnt[tid];
                if (v) atomicAdd(&p.counts[(size_t)sec_base * n_pix + roach * p.npix_per_roach + tid], v);
            }
            if (tid < 256) cnt[tid] = 0;                      // reused two chunks later (a barrier lies in between)
            if (SMEM_HIST && p.hist) {
                const int hs = p.n_bins + 1, n = p.npix_per_roach * hs;
                if (tid < p.npix_per_roach) {                 // per-pixel count of this chunk = row sum (incl. overflow column)
                    uint32_t sum = 0;
                    for (int b = 0; b < hs; ++b) sum += s_hist[tid * hs + b];
                    if (sum && sec_base < p.exptime)
                        atomicAdd(&p.counts[(size_t)sec_base * n_pix + roach * p.npix_per_roach + tid], sum);
                }
                workers_sync();
                for (int i = tid; i < n; i += DEC_WORKERS) {
                    const uint32_t v = s_hist[i];
                    if (v) {
                        const int pix = i / hs, b = i - pix * hs;
                        if (b < p.n_bins) atomicAdd(&p.hist[((size_t)roach * p.npix_per_roach + pix) * p.n_bins + b], v);
                        s_hist[i] = 0;
                    }
                }
                workers_sync();                               // s_hist is shared by consecutive chunks
            }
            if (p.prof && tid == 0) {
                const long long w2 = clock64();
                atomicAdd(&p.prof[5], (unsigned long long)(w1 - w0));   // workers: wait for a ready chunk
                atomicAdd(&p.prof[6], (unsigned long long)(w2 - w1));   // workers: process + flush
            }
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            n_eos_t += __shfl_xor_sync(0xffffffffu, n_eos_t, d);
            n_bad += __shfl_xor_sync(0xffffffffu, n_bad, d);
            n_nonpix += __shfl_xor_sync(0xffffffffu, n_nonpix, d);
            n_ign += __shfl_xor_sync(0xffffffffu, n_ign, d);
            n_ok += __shfl_xor_sync(0xffffffffu, n_ok, d);
        }
        if (lane == 0) {
            if (n_eos_t) atomicAdd(&s_stat[0], (unsigned long long)n_eos_t);
            if (n_bad) atomicAdd(&s_stat[1], (unsigned long long)n_bad);
            if (n_nonpix) atomicAdd(&s_stat[2], (unsigned long long)n_nonpix);
            if (n_ign) atomicAdd(&s_stat[3], (unsigned long long)n_ign);
            if (n_ok) atomicAdd(&s_stat[4], (unsigned long long)n_ok);
        }
    }
    __syncthreads();
    if (tid < 5 && s_stat[tid]) atomicAdd(&p.stats[tid], s_stat[tid]);
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

int decode_common(mkid_ctx *ctx, const uint64_t *words, const uint32_t *wire, int64_t n_units,
                  const int64_t *seg_offset, const int64_t *seg_len_in, const int32_t *seg_roach, const int32_t *seg_sec,
                  int32_t *seg_sec_out, int32_t n_seg, const mkid_decode_cfg *cfg, uint32_t *counts_raw,
                  uint32_t *hist, mkid_decode_stats *stats) {
    MKID_REQUIRE(ctx, cfg && seg_offset && seg_roach && n_seg > 0, "decode: missing cfg/segments");
    MKID_REQUIRE(ctx, cfg->npix_per_roach > 0 && cfg->npix_per_roach <= 255, "npix_per_roach must be 1..255");
    MKID_REQUIRE(ctx, cfg->n_roaches > 0 && cfg->exptime > 0 && counts_raw, "bad decode cfg");
    const bool want_hist = hist != nullptr && cfg->hist_field_shift >= 0;
    if (want_hist) MKID_REQUIRE(ctx, cfg->n_bins > 0 && cfg->hist_field_shift <= 52, "bad histogram cfg");
    MKID_CUDA(ctx, cudaSetDevice(ctx->device));
    const bool wire_fmt = wire != nullptr;
    const int64_t n_pix = (int64_t)cfg->n_roaches * cfg->npix_per_roach;

    // segment table (host) -> chunk prefix
    std::vector<int64_t> first_chunk(n_seg + 1, 0);
    std::vector<int64_t> seg_len(n_seg);
    for (int i = 0; i < n_seg; ++i) {
        int64_t len = seg_len_in ? seg_len_in[i] : seg_offset[i + 1] - seg_offset[i];
        seg_len[i] = len;
        MKID_REQUIRE(ctx, len >= 0 && seg_offset[i] >= 0 && seg_offset[i] + len <= n_units, "segment offsets out of range");
        MKID_REQUIRE(ctx, seg_roach[i] >= 0 && seg_roach[i] < cfg->n_roaches, "segment roach out of range");
        first_chunk[i + 1] = first_chunk[i] + (wire_fmt ? 2 * len : (len + DEC_CHUNK - 1) / DEC_CHUNK);
    }
    const int64_t n_chunks = first_chunk[n_seg];
    std::vector<int32_t> sec0(n_seg, 0);
    if (seg_sec) for (int i = 0; i < n_seg; ++i) sec0[i] = seg_sec[i];

    // meta buffer: first_chunk | seg_offset | seg_len | stats(5 u64) | roach | sec | sec_out | ticket
    const size_t meta_bytes = (size_t)(n_seg + 1) * 24 + (size_t)n_seg * 12 + 5 * 8 + 16;
    char *meta = nullptr;
    int rc = mkid_scratch(ctx, SCR_META, meta_bytes, (void **)&meta);
    if (rc) return rc;
    int64_t *d_first = (int64_t *)meta;
    int64_t *d_off = d_first + (n_seg + 1);
    int64_t *d_len = d_off + (n_seg + 1);
    unsigned long long *d_stats = (unsigned long long *)(d_len + (n_seg + 1));
    int32_t *d_roach = (int32_t *)(d_stats + 5);
    int32_t *d_sec = d_roach + n_seg;
    int32_t *d_sec_out = d_sec + n_seg;
    unsigned int *d_ticket = (unsigned int *)(d_sec_out + n_seg);
    {   // upload the segment table only when it changed since the last call on this context
        std::vector<char> blob(meta_bytes, 0);
        memcpy(blob.data() + ((char *)d_first - meta), first_chunk.data(), (n_seg + 1) * 8);
        memcpy(blob.data() + ((char *)d_off - meta), seg_offset, n_seg * 8);
        memcpy(blob.data() + ((char *)d_len - meta), seg_len.data(), n_seg * 8);
        memcpy(blob.data() + ((char *)d_roach - meta), seg_roach, n_seg * 4);
        memcpy(blob.data() + ((char *)d_sec - meta), sec0.data(), n_seg * 4);
        if (ctx->dec_meta_dev != meta || ctx->dec_meta_host != blob) {
            MKID_CUDA(ctx, cudaMemcpyAsync(meta, blob.data(), meta_bytes, cudaMemcpyHostToDevice, ctx->stream));
            MKID_CUDA(ctx, cudaStreamSynchronize(ctx->stream));   // blob is pageable and about to be moved
            ctx->dec_meta_host.swap(blob);
            ctx->dec_meta_dev = meta;
        } else {
            MKID_CUDA(ctx, cudaMemsetAsync(d_stats, 0, 5 * 8, ctx->stream));
            MKID_CUDA(ctx, cudaMemsetAsync(d_sec_out, 0, n_seg * 4 + 16, ctx->stream));
        }
    }

    unsigned long long *d_state = nullptr;
    rc = mkid_scratch(ctx, SCR_STATE, (size_t)(n_chunks + 1) * 8, (void **)&d_state);
    if (rc) return rc;
    MKID_CUDA(ctx, cudaMemsetAsync(d_state, 0, (size_t)(n_chunks + 1) * 8, ctx->stream));

    const void *d_in = nullptr;
    const size_t in_bytes = wire_fmt ? (size_t)n_units * 2 * DEC_BUNDLE * 4 : (size_t)n_units * 8;
    rc = mkid_stage_in(ctx, wire_fmt ? (const void *)wire : (const void *)words, in_bytes, SCR_IN, &d_in);
    if (rc) return rc;
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

    DecParams p;
    p.words = wire_fmt ? nullptr : (const uint64_t *)d_in;
    p.wire = wire_fmt ? (const uint32_t *)d_in : nullptr;
    p.seg_first_chunk = d_first; p.seg_offset = d_off; p.seg_len = d_len; p.seg_roach = d_roach; p.seg_sec = d_sec;
    p.seg_sec_out = d_sec_out; p.n_seg = n_seg; p.n_chunks = n_chunks;
    p.n_roaches = cfg->n_roaches; p.npix_per_roach = cfg->npix_per_roach; p.exptime = cfg->exptime;
    p.field_shift = want_hist ? cfg->hist_field_shift : 0; p.n_bins = want_hist ? cfg->n_bins : 0;
    p.bin_lut = (const uint16_t *)d_lut; p.counts = (uint32_t *)d_counts; p.hist = (uint32_t *)d_hist;
    p.stats = d_stats; p.state = d_state; p.ticket = d_ticket;
    p.prof = nullptr;
    static const bool want_prof = getenv("MKID_DEC_PROFILE") != nullptr;
    if (want_prof) {
        void *pb = nullptr;
        if ((rc = mkid_scratch(ctx, SCR_AUX5, 64, &pb))) return rc;
        MKID_CUDA(ctx, cudaMemsetAsync(pb, 0, 64, ctx->stream));
        p.prof = (unsigned long long *)pb;
    }

    if (n_chunks > 0) {
        const bool smem_hist = want_hist && (int64_t)cfg->npix_per_roach * (cfg->n_bins + 1) <= DEC_SMEM_HIST;
        int grid = (int)std::min<int64_t>(n_chunks, (int64_t)ctx->num_sms);   // persistent: one CTA per SM
        const size_t dyn = (size_t)DEC_STAGES * DEC_STAGE_BYTES;
        auto launch = [&](auto kern) -> cudaError_t {
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
            if (e != cudaSuccess) return e;
            kern<<<grid, DEC_THREADS, dyn, ctx->stream>>>(p);
            return cudaSuccess;
        };
        cudaError_t le;
        if (wire_fmt) le = smem_hist ? launch(decode_kernel<true, true>) : launch(decode_kernel<true, false>);
        else le = smem_hist ? launch(decode_kernel<false, true>) : launch(decode_kernel<false, false>);
        MKID_CUDA(ctx, le);
        MKID_CHECK_LAUNCH(ctx);
    }
    if (p.prof) {
        unsigned long long h[8];
        MKID_CUDA(ctx, cudaMemcpyAsync(h, p.prof, 64, cudaMemcpyDeviceToHost, ctx->stream));
        MKID_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        const double n = h[7] ? (double)h[7] : 1.0;
        fprintf(stderr, "[mkid decode profile] chunks %llu | scout cycles/chunk: wait-free %.0f ticket %.0f load %.0f scan %.0f "
                        "lookback %.0f | workers cycles/chunk: wait-ready %.0f process %.0f\n", h[7], h[0] / n, h[1] / n,
                h[2] / n, h[3] / n, h[4] / n, h[5] / n, h[6] / n);
    }
    rc = mkid_stage_out_finish(ctx, counts_raw, counts_bytes, d_counts);
    if (rc) return rc;
    if (want_hist) {
        rc = mkid_stage_out_finish(ctx, hist, hist_bytes, d_hist);
        if (rc) return rc;
    }
    if (seg_sec_out) {
        if (n_chunks == 0) { for (int i = 0; i < n_seg; ++i) seg_sec_out[i] = sec0[i]; }
        else {
            std::vector<int32_t> tmp(n_seg);
            MKID_CUDA(ctx, cudaMemcpyAsync(tmp.data(), d_sec_out, n_seg * 4, cudaMemcpyDeviceToHost, ctx->stream));
            MKID_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            for (int i = 0; i < n_seg; ++i)
                seg_sec_out[i] = (first_chunk[i + 1] == first_chunk[i]) ? sec0[i] : tmp[i];
        }
    }
    if (stats) {
        if (mkid_is_device_ptr(stats)) {
            unsigned long long h[5];
            MKID_CUDA(ctx, cudaMemcpyAsync(h, d_stats, 40, cudaMemcpyDeviceToHost, ctx->stream));
            MKID_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            mkid_decode_stats cur;
            MKID_CUDA(ctx, cudaMemcpy(&cur, stats, sizeof(cur), cudaMemcpyDeviceToHost));
            cur.n_eos += h[0]; cur.n_corrupt_eos += h[1]; cur.n_nonpixel += h[2]; cur.n_ignored += h[3]; cur.n_valid += h[4];
            MKID_CUDA(ctx, cudaMemcpy(stats, &cur, sizeof(cur), cudaMemcpyHostToDevice));
        } else {
            unsigned long long h[5];
            MKID_CUDA(ctx, cudaMemcpyAsync(h, d_stats, 40, cudaMemcpyDeviceToHost, ctx->stream));
            MKID_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
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
