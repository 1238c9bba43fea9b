// K6: photon-word decode + per-(second,pixel) binning + per-pixel pulse-height histogram.
//
// Replaces the receive/bin loop of DataReadout/ReadoutControls/lib/PacketMaster.c:286-397 and
// the per-word unpack of DataReadout/ChannelizerControls/ROACH_Pulses.py:795-832.
//
// HBM-bound integer work: every word is read exactly once (8 B/word).  A chunk is 8192 words
// (= one PulseServer bundle, PacketMaster.c:42-44).  Persistent CTAs take chunks from an
// ordered ticket; which second a word belongs to is the number of end-of-second words before
// it in its roach stream, resolved in the same pass with a decoupled look-back over per-chunk
// EOS counts (chunk c waits only for the published prefix of chunk c-1 of its segment).
#include "common.cuh"

namespace {

constexpr int DEC_THREADS = 512;
constexpr int DEC_CHUNK = 4096;                    // words per chunk (wire: half a bundle)
constexpr int DEC_BUNDLE = 8192;                   // PacketMaster.c:44 BUFSIZE_INTS
constexpr int DEC_WPT = DEC_CHUNK / DEC_THREADS;   // 8 words per thread
constexpr int DEC_ITERS = DEC_WPT / 2;             // 4 iterations of 2 words (one uint4)
constexpr int DEC_WARPS = DEC_THREADS / 32;
constexpr int DEC_SMEM_HIST = 4096;                // smem-privatised histogram entries

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
};

__device__ __forceinline__ uint32_t bswap32(uint32_t x) { return __byte_perm(x, 0, 0x0123); }

template <bool WIRE, bool SMEM_HIST>
__global__ void __launch_bounds__(DEC_THREADS, 2) decode_kernel(DecParams p) {
    __shared__ uint32_t s_cnt[256];
    __shared__ uint32_t s_hist[SMEM_HIST ? DEC_SMEM_HIST : 1];
    __shared__ uint16_t s_lut[4096];
    __shared__ int s_eos_iw[DEC_ITERS][DEC_WARPS];
    __shared__ unsigned long long s_stat[5];
    __shared__ long long s_chunk;
    __shared__ int s_seg, s_secbase;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool use_lut = p.bin_lut != nullptr;
    if (use_lut)
        for (int i = tid; i < 4096; i += DEC_THREADS) s_lut[i] = p.bin_lut[i];
    if (tid < 5) s_stat[tid] = 0;
    const int n_pix = p.n_roaches * p.npix_per_roach;

    for (;;) {
        __syncthreads();   // previous iteration's smem fully consumed
        if (tid == 0) {
            long long c = (long long)atomicAdd(p.ticket, 1u);
            s_chunk = c;
            if (c < p.n_chunks) {   // segment of this chunk: last g with first_chunk[g] <= c
                int lo = 0, hi = p.n_seg;
                while (hi - lo > 1) {
                    int mid = (lo + hi) >> 1;
                    if (p.seg_first_chunk[mid] <= c) lo = mid; else hi = mid;
                }
                s_seg = lo;
            }
        }
        if (tid < 256) s_cnt[tid] = 0;
        if (SMEM_HIST)
            for (int i = tid; i < DEC_SMEM_HIST; i += DEC_THREADS) s_hist[i] = 0;
        __syncthreads();
        const long long c = s_chunk;
        if (c >= p.n_chunks) break;
        const int g = s_seg;
        const long long lc = c - p.seg_first_chunk[g];
        const bool last_chunk = (c + 1 == p.seg_first_chunk[g + 1]);
        const int roach = p.seg_roach[g];

        // ---- load the chunk: each thread 8 x (2 words), coalesced 16 B per lane
        uint64_t w[DEC_WPT];
        int n_here;   // valid words in this chunk
        if (WIRE) {
            n_here = DEC_CHUNK;
            // chunk lc of a segment = half (lc & 1) of bundle lc >> 1: 4096 low halves + 4096 high halves
            const uint32_t *lo_blk = p.wire + (size_t)(p.seg_offset[g] + (lc >> 1)) * (2 * DEC_BUNDLE) + (lc & 1) * DEC_CHUNK;
            const uint32_t *hi_blk = lo_blk + DEC_BUNDLE;
#pragma unroll
            for (int i = 0; i < DEC_ITERS / 2; ++i) {     // 2 iterations of 4 words
                uint4 l = ld_stream_u4(reinterpret_cast<const uint4 *>(lo_blk) + i * DEC_THREADS + tid);
                uint4 h = ld_stream_u4(reinterpret_cast<const uint4 *>(hi_blk) + i * DEC_THREADS + tid);
                w[4 * i + 0] = ((uint64_t)bswap32(h.x) << 32) | bswap32(l.x);
                w[4 * i + 1] = ((uint64_t)bswap32(h.y) << 32) | bswap32(l.y);
                w[4 * i + 2] = ((uint64_t)bswap32(h.z) << 32) | bswap32(l.z);
                w[4 * i + 3] = ((uint64_t)bswap32(h.w) << 32) | bswap32(l.w);
            }
        } else {
            const long long seg_n = p.seg_len[g];
            const long long rem = seg_n - lc * DEC_CHUNK;
            n_here = rem < DEC_CHUNK ? (int)rem : DEC_CHUNK;
            const uint64_t *base = p.words + p.seg_offset[g] + lc * DEC_CHUNK;
            if (n_here == DEC_CHUNK && ((reinterpret_cast<uintptr_t>(base) & 15) == 0)) {
#pragma unroll
                for (int i = 0; i < DEC_ITERS; ++i) {
                    uint4 v = ld_stream_u4(reinterpret_cast<const uint4 *>(base) + i * DEC_THREADS + tid);
                    w[2 * i + 0] = ((uint64_t)v.y << 32) | v.x;
                    w[2 * i + 1] = ((uint64_t)v.w << 32) | v.z;
                }
            } else {   // ragged tail or odd alignment: scalar loads, same position mapping
#pragma unroll
                for (int i = 0; i < DEC_ITERS; ++i) {
                    int pos = (i * DEC_THREADS + tid) * 2;
                    w[2 * i + 0] = pos < n_here ? base[pos] : 0ull;
                    w[2 * i + 1] = pos + 1 < n_here ? base[pos + 1] : 0ull;
                }
            }
        }
        // position of w[k] inside the chunk
        auto pos_of = [&](int k) -> int {
            if (WIRE) return ((k >> 2) * DEC_THREADS + tid) * 4 + (k & 3);
            return ((k >> 1) * DEC_THREADS + tid) * 2 + (k & 1);
        };

        // ---- end-of-second words in this chunk (rare): ordered local rank of every word
        uint32_t eos_bits = 0;
#pragma unroll
        for (int k = 0; k < DEC_WPT; ++k)
            if ((uint32_t)(w[k] >> 56) == 255u && pos_of(k) < n_here) eos_bits |= 1u << k;
        const int any_eos = __syncthreads_or(eos_bits != 0);
        constexpr int G = WIRE ? DEC_ITERS / 2 : DEC_ITERS;    // load groups per thread
        constexpr int S = WIRE ? 4 : 2;                         // words per group
        int rbase[G];        // seconds closed inside the chunk before the first word of group gi
        int total_eos = 0;
#pragma unroll
        for (int gi = 0; gi < G; ++gi) rbase[gi] = 0;
        if (any_eos) {
            // positions ascend with (group, tid, sub): per group a warp-ordered exclusive prefix
#pragma unroll
            for (int gi = 0; gi < G; ++gi) {
                const int mine = __popc((eos_bits >> (gi * S)) & ((1u << S) - 1));
                int incl = mine;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    int t = __shfl_up_sync(0xffffffffu, incl, d);
                    if (lane >= d) incl += t;
                }
                rbase[gi] = incl - mine;
                if (lane == 31) s_eos_iw[gi][warp] = incl;
            }
            __syncthreads();
            int run = 0;   // every thread scans the (G x WARPS) table (<= 128 entries)
#pragma unroll
            for (int gi = 0; gi < G; ++gi) {
                for (int ww = 0; ww < DEC_WARPS; ++ww) {
                    if (ww == warp) rbase[gi] += run;
                    run += s_eos_iw[gi][ww];
                }
            }
            total_eos = run;
        }

        // ---- decoupled look-back (warp-parallel): seconds closed before this chunk
        if (warp == 0) {
            int base_sec = 0;
            if (lc == 0) {
                base_sec = p.seg_sec[g];
            } else {
                // publish the aggregate first so successors never wait on our own look-back
                if (lane == 0) atomicExch(&p.state[c], (1ull << 32) | (unsigned)total_eos);
                const long long first = c - lc;       // first chunk of the segment: always inclusive
                long long q = c - 1;
                int acc = 0;
                for (;;) {
                    const long long idx = q - lane;
                    const bool valid = idx >= first;
                    unsigned long long st = 0;
                    if (valid) {
                        do { st = *reinterpret_cast<volatile unsigned long long *>(&p.state[idx]); } while ((st >> 32) == 0);
                    }
                    const unsigned incl = __ballot_sync(0xffffffffu, valid && (st >> 32) == 2);
                    int v;
                    if (incl) {
                        const int stop = __ffs(incl) - 1;           // nearest predecessor holding a prefix
                        v = (lane <= stop) ? (int)(unsigned)st : 0;
                    } else {
                        v = valid ? (int)(unsigned)st : 0;
                    }
#pragma unroll
                    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
                    acc += v;
                    if (incl) break;
                    q -= 32;
                }
                base_sec = acc;
            }
            if (lane == 0) {
                // the record is self-contained (flag and value in one 64-bit word): no fence needed
                atomicExch(&p.state[c], (2ull << 32) | (unsigned)(base_sec + total_eos));
                s_secbase = base_sec;
                if (last_chunk && p.seg_sec_out) p.seg_sec_out[g] = base_sec + total_eos;
            }
        }
        __syncthreads();
        const int sec_base = s_secbase;

        // ---- bin
        unsigned n_eos = 0, n_bad = 0, n_nonpix = 0, n_ign = 0, n_ok = 0;
#pragma unroll
        for (int k = 0; k < DEC_WPT; ++k) {
            if (pos_of(k) >= n_here) continue;
            const uint64_t x = w[k];
            const int l = rbase[k / S] + __popc((eos_bits >> ((k / S) * S)) & ((1u << (k % S)) - 1));
            const int sec = sec_base + l;
            const uint32_t adr = (uint32_t)(x >> 56);
            if (sec >= p.exptime) { ++n_ign; continue; }
            if (adr == 255u) {
                ++n_eos;
                if (x != ~0ull) ++n_bad;
                continue;
            }
            if ((int)adr >= p.npix_per_roach) { ++n_nonpix; continue; }
            ++n_ok;
            if (l == 0) atomicAdd(&s_cnt[adr], 1u);
            else atomicAdd(&p.counts[(size_t)sec * n_pix + roach * p.npix_per_roach + adr], 1u);
            if (p.hist) {
                uint32_t f = (uint32_t)(x >> p.field_shift) & 0xFFFu;
                uint32_t b = use_lut ? s_lut[f] : f;
                if ((int)b < p.n_bins) {
                    if (SMEM_HIST) atomicAdd(&s_hist[adr * p.n_bins + b], 1u);
                    else atomicAdd(&p.hist[((size_t)(roach * p.npix_per_roach + adr)) * p.n_bins + b], 1u);
                }
            }
        }
        // stats: warp reduce then one smem atomic per warp
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            n_eos += __shfl_xor_sync(0xffffffffu, n_eos, d);
            n_bad += __shfl_xor_sync(0xffffffffu, n_bad, d);
            n_nonpix += __shfl_xor_sync(0xffffffffu, n_nonpix, d);
            n_ign += __shfl_xor_sync(0xffffffffu, n_ign, d);
            n_ok += __shfl_xor_sync(0xffffffffu, n_ok, d);
        }
        if (lane == 0) {
            if (n_eos) atomicAdd(&s_stat[0], (unsigned long long)n_eos);
            if (n_bad) atomicAdd(&s_stat[1], (unsigned long long)n_bad);
            if (n_nonpix) atomicAdd(&s_stat[2], (unsigned long long)n_nonpix);
            if (n_ign) atomicAdd(&s_stat[3], (unsigned long long)n_ign);
            if (n_ok) atomicAdd(&s_stat[4], (unsigned long long)n_ok);
        }
        __syncthreads();
        // ---- flush the privatised counters of this chunk
        if (sec_base < p.exptime && tid < p.npix_per_roach && tid < 256) {
            uint32_t v = s_cnt[tid];
            if (v) atomicAdd(&p.counts[(size_t)sec_base * n_pix + roach * p.npix_per_roach + tid], v);
        }
        if (SMEM_HIST && p.hist) {
            const int n = p.npix_per_roach * p.n_bins;
            for (int i = tid; i < n; i += DEC_THREADS) {
                uint32_t v = s_hist[i];
                if (v) atomicAdd(&p.hist[(size_t)roach * n + i], v);
            }
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

    if (n_chunks > 0) {
        const bool smem_hist = want_hist && (int64_t)cfg->npix_per_roach * cfg->n_bins <= DEC_SMEM_HIST;
        int grid = (int)std::min<int64_t>(n_chunks, (int64_t)ctx->num_sms * 2);   // persistent, 2 CTAs of 512 threads per SM
        if (wire_fmt) {
            if (smem_hist) decode_kernel<true, true><<<grid, DEC_THREADS, 0, ctx->stream>>>(p);
            else decode_kernel<true, false><<<grid, DEC_THREADS, 0, ctx->stream>>>(p);
        } else {
            if (smem_hist) decode_kernel<false, true><<<grid, DEC_THREADS, 0, ctx->stream>>>(p);
            else decode_kernel<false, false><<<grid, DEC_THREADS, 0, ctx->stream>>>(p);
        }
        MKID_CHECK_LAUNCH(ctx);
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
