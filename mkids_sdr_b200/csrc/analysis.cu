// Analysis kernels next to the hot path: the reference's snapshot decoders, its
// float64 software pulse triggers and the threshold derivation of loadThresholds, on the GPU.
//
//   mkid_iq_snapshot_decode   pulse_triggering_IQ.py:121-147 (40-bit I/Q snapshot words)
//   mkid_phase_deg_from_iq    pulse_triggering_IQ.py:152
//   mkid_soft_trigger         pulse_triggering_v2.py:104-174 (rolling mean), pulse_triggering.py:114-208 and
//                             contsnapshot ROACH_Pulses.py:614-725 (block mean)
//   mkid_thresholds_from_phase  ROACH_Pulses.py:259-288 (np.histogram(bins=100) -> CDF -> median / 5 % edge)
//
// Everything that decides a comparison is float64 in the operation order of the NumPy calls the reference makes
// (np.mean = pairwise summation with 8 accumulators, np.histogram = linspace edges + edge corrections), so hit
// lists and thresholds are bit-identical to the oracle (oracle/trigger.py, oracle/control.py).
#include <math.h>

#include "common.cuh"

namespace {

// ---------------------------------------------------------------- a8 / a9
__global__ void iq_snapshot_kernel(const uint8_t *__restrict__ buf, int64_t n_words16, int16_t *I, int16_t *Q) {
    int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; k < n_words16; k += stride) {
        const uint4 v = reinterpret_cast<const uint4 *>(buf)[k];
        uint8_t b[16];
        memcpy(b, &v, 16);
        // bytes 6-10 and 11-15: 20-bit I field (low 16 bits kept) then 16-bit Q, big-endian
        const uint32_t i0 = ((uint32_t)(b[6] & 0xF) << 12) | ((uint32_t)b[7] << 4) | (b[8] >> 4);
        const uint32_t q0 = ((uint32_t)b[9] << 8) | b[10];
        const uint32_t i1 = ((uint32_t)(b[11] & 0xF) << 12) | ((uint32_t)b[12] << 4) | (b[13] >> 4);
        const uint32_t q1 = ((uint32_t)b[14] << 8) | b[15];
        I[2 * k] = (int16_t)i0; I[2 * k + 1] = (int16_t)i1;       // twos_comp(., 16), pulse_triggering.py:22-26
        Q[2 * k] = (int16_t)q0; Q[2 * k + 1] = (int16_t)q1;
    }
}

__global__ void phase_deg_kernel(const int16_t *__restrict__ I, const int16_t *__restrict__ Q, int64_t n, double Ic,
                                 double Qc, double *deg) {
    int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const double two_pi = 2.0 * 3.141592653589793;
    for (; k < n; k += stride) {
        const double a = atan2((double)Q[k] - Qc, (double)I[k] - Ic);
        deg[k] = __ddiv_rn(__dmul_rn(-360.0, a), two_pi);           // -360*(arctan2(..))/(2*pi)
    }
}

// ---------------------------------------------------------------- np.add.reduce on float64 (pairwise, numpy >= 1.9)
template <typename LOAD>
__device__ double pairwise_sum(LOAD ld, int64_t off, int64_t n) {
    if (n < 8) {
        double res = 0.0;
        for (int64_t i = 0; i < n; ++i) res = __dadd_rn(res, ld(off + i));
        return res;
    }
    if (n <= 128) {
        double r[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) r[j] = ld(off + j);
        int64_t i = 8;
        for (; i < n - (n % 8); i += 8) {
#pragma unroll
            for (int j = 0; j < 8; ++j) r[j] = __dadd_rn(r[j], ld(off + i + j));
        }
        double res = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])),
                               __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
        for (; i < n; ++i) res = __dadd_rn(res, ld(off + i));
        return res;
    }
    int64_t n2 = n / 2;
    n2 -= n2 % 8;
    return __dadd_rn(pairwise_sum(ld, off, n2), pairwise_sum(ld, off + n2, n - n2));
}
template <typename LOAD>
__device__ double sequential_sum(LOAD ld, int64_t off, int64_t n) {
    double res = 0.0;
    for (int64_t i = 0; i < n; ++i) res = __dadd_rn(res, ld(off + i));
    return res;
}

struct TrigParams {
    const double *x;          // [n_streams][n]
    int64_t n;
    int n_streams;
    int mode, M, wrap, seq;
    double thr;
    double *means;            // block mode: [n_streams][n / M]
    uint32_t *mask;           // [n_streams][n_mask_words]
    int64_t n_mask_words;
};

__device__ __forceinline__ double wrapped(const double *x, int64_t i, int wrap) {
    const double v = x[i];
    return (wrap && v < 0.0) ? __dadd_rn(v, 360.0) : v;             // pulse_triggering.py:110-112
}

__global__ void trig_block_means_kernel(TrigParams p) {
    const int s = blockIdx.y;
    const int64_t n_means = p.n / p.M;
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_means) return;
    const double *x = p.x + (size_t)s * p.n;
    const int wrap = p.wrap;
    auto ld = [&](int64_t i) { return wrapped(x, i, wrap); };
    const double sum = p.seq ? sequential_sum(ld, j * p.M, p.M) : pairwise_sum(ld, j * p.M, p.M);
    p.means[(size_t)s * n_means + j] = __ddiv_rn(sum, (double)p.M);   // np.mean = add.reduce / count
}

__global__ void trig_candidates_kernel(TrigParams p) {
    const int s = blockIdx.y;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const double *x = p.x + (size_t)s * p.n;
    bool cand = false;
    if (t < p.n) {
        if (p.mode == 0) {
            if (t >= p.M) {
                auto ld = [&](int64_t i) { return x[i]; };
                const double sum = p.seq ? sequential_sum(ld, t - p.M, p.M) : pairwise_sum(ld, t - p.M, p.M);
                const double mean = __ddiv_rn(sum, (double)p.M);
                cand = fabs(__dsub_rn(mean, x[t])) > p.thr;           // pulse_triggering_v2.py:115-119
            }
        } else {
            const int64_t wm = t / p.M, n_means = p.n / p.M;
            if (wm < n_means) cand = fabs(__dsub_rn(p.means[(size_t)s * n_means + wm], wrapped(x, t, p.wrap))) > p.thr;
        }
    }
    const unsigned bal = __ballot_sync(0xffffffffu, cand);
    if ((threadIdx.x & 31) == 0 && (t >> 5) < p.n_mask_words) p.mask[(size_t)s * p.n_mask_words + (t >> 5)] = bal;
}

// greedy hold-off: one warp per stream
__global__ void trig_greedy_kernel(const uint32_t *__restrict__ mask, int64_t n_mask_words, int64_t n, int n_streams,
                                   int64_t start, int64_t holdoff, int64_t tail, int32_t *hits, int max_hits,
                                   int32_t *n_hits) {
    const int s = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (s >= n_streams) return;
    const uint32_t *mk = mask + (size_t)s * n_mask_words;
    int32_t *out = hits + (size_t)s * max_hits;
    int64_t bob = start;
    int count = 0;
    while (bob < n) {
        const int64_t w0 = bob >> 5;
        const int64_t wi = w0 + lane;
        uint32_t w = wi < n_mask_words ? mk[wi] : 0u;
        if (lane == 0) w &= 0xFFFFFFFFu << (bob & 31);
        const unsigned nz = __ballot_sync(0xffffffffu, w != 0u);
        if (!nz) { bob = (w0 + 32) << 5; continue; }
        const int src = __ffs(nz) - 1;
        const uint32_t ww = __shfl_sync(0xffffffffu, w, src);
        const int64_t t = ((w0 + src) << 5) + (__ffs(ww) - 1);
        if (t + tail > n) break;                                      // the literal loops' break test
        if (lane == 0 && count < max_hits) out[count] = (int32_t)t;
        ++count;
        bob = t + holdoff;
    }
    if (lane == 0) n_hits[s] = count;
}

// ---------------------------------------------------------------- a10: loadThresholds
// One CTA per channel.  phase element (t, c) of board b at phase[b*board_stride + t*row_stride + c].
constexpr int TH_BINS = 100;
__global__ void __launch_bounds__(256) thresholds_kernel(const int16_t *__restrict__ phase, int64_t board_stride,
                                                         int64_t row_stride, int n_ch, int64_t n, double nsigma,
                                                         int32_t floor_raw, int32_t *thr_raw, double *med_out, double *p5_out) {
    __shared__ int s_min, s_max;
    __shared__ unsigned s_hist[TH_BINS];
    __shared__ double s_edges[TH_BINS + 1];
    __shared__ double s_nrm[TH_BINS];
    const int c = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
    const int16_t *x = phase + (size_t)b * board_stride + c;
    if (tid == 0) { s_min = 32767; s_max = -32768; }
    if (tid < TH_BINS) s_hist[tid] = 0;
    __syncthreads();
    int mn = 32767, mx = -32768;
    for (int64_t t = tid; t < n; t += 256) { const int v = x[t * row_stride]; mn = min(mn, v); mx = max(mx, v); }
    mn = __reduce_min_sync(0xffffffffu, mn); mx = __reduce_max_sync(0xffffffffu, mx);
    if ((tid & 31) == 0) { atomicMin(&s_min, mn); atomicMax(&s_max, mx); }
    __syncthreads();
    // np.histogram(a, bins=100): outer edges (min, max), widened by 0.5 when equal; edges = linspace
    double first = (double)s_min, last = (double)s_max;
    const bool flat = s_min == s_max;
    if (flat) { first -= 0.5; last += 0.5; }
    const double delta = __dsub_rn(last, first), step = __ddiv_rn(delta, (double)TH_BINS);
    if (tid <= TH_BINS) s_edges[tid] = tid == TH_BINS ? last : __dadd_rn(__dmul_rn((double)tid, step), first);
    __syncthreads();
    // bin = int(((a - first_edge) / (last_edge - first_edge)) * 100), == 100 -> 99, then the two edge corrections
    const double denom = flat ? 1.0 : (double)(unsigned)(s_max - s_min);
    for (int64_t t = tid; t < n; t += 256) {
        const int v = x[t * row_stride];
        const double a = (double)v;
        const double num = flat ? __dsub_rn(a, first) : (double)(unsigned)(v - s_min);
        int idx = (int)__dmul_rn(__ddiv_rn(num, denom), (double)TH_BINS);
        if (idx == TH_BINS) idx = TH_BINS - 1;
        if (a < s_edges[idx]) --idx;
        else if (a >= s_edges[idx + 1] && idx != TH_BINS - 1) ++idx;
        atomicAdd(&s_hist[idx], 1u);
    }
    __syncthreads();
    // n = float32(counts) / sum (float64 division), tot[i] = np.sum(n[:i]) (pairwise), argmin |tot - 0.5|, |tot - 0.05|
    if (tid < TH_BINS) s_nrm[tid] = __ddiv_rn((double)(float)s_hist[tid], (double)n);
    __syncthreads();
    double d50 = 1e300, d05 = 1e300;
    int i50 = 0, i05 = 0;
    for (int i = tid; i <= TH_BINS; i += 256) {
        auto ld = [&](int64_t k) { return s_nrm[k]; };
        const double tot = pairwise_sum(ld, 0, i);
        d50 = fabs(__dsub_rn(tot, 0.5)); i50 = i;
        d05 = fabs(__dsub_rn(tot, 0.05)); i05 = i;
    }
    // first minimal index: reduce (distance, index) lexicographically over the 101 threads that hold one
    __shared__ double r_d50[TH_BINS + 1], r_d05[TH_BINS + 1];
    if (tid <= TH_BINS) { r_d50[tid] = d50; r_d05[tid] = d05; }
    __syncthreads();
    if (tid == 0) {
        int a50 = 0, a05 = 0;
        for (int i = 1; i <= TH_BINS; ++i) {
            if (r_d50[i] < r_d50[a50]) a50 = i;
            if (r_d05[i] < r_d05[a05]) a05 = i;
        }
        const double med = s_edges[a50], p5 = s_edges[a05];
        int thr = (int)__dmul_rn(-nsigma, fabs(__dsub_rn(med, p5)));      // int() truncates toward zero
        if (thr < floor_raw) thr = floor_raw;
        const size_t o = (size_t)b * n_ch + c;
        thr_raw[o] = thr;
        if (med_out) med_out[o] = med;
        if (p5_out) p5_out[o] = p5;
    }
    (void)i50; (void)i05;
}

}  // namespace

extern "C" int mkid_iq_snapshot_decode(mkid_ctx *ctx, const uint8_t *buf, int64_t n_bytes, int16_t *I, int16_t *Q) {
    if (!ctx) return MKID_EINVAL;
    MKID_REQUIRE(ctx, buf && I && Q && n_bytes >= 0 && n_bytes % 16 == 0, "iq_snapshot_decode: need whole 16-byte words");
    if (n_bytes == 0) return MKID_OK;
    MKID_CUDA(ctx, cudaSetDevice(ctx->device));
    const int64_t nw = n_bytes / 16;
    const void *d_in; void *d_i, *d_q; int rc;
    if ((rc = mkid_stage_in(ctx, buf, n_bytes, SCR_IN, &d_in))) return rc;
    MKID_REQUIRE(ctx, (reinterpret_cast<uintptr_t>(d_in) & 15) == 0, "iq_snapshot_decode: buffer must be 16-byte aligned");
    if ((rc = mkid_stage_out(ctx, I, nw * 4, SCR_OUT0, false, &d_i))) return rc;
    if ((rc = mkid_stage_out(ctx, Q, nw * 4, SCR_OUT1, false, &d_q))) return rc;
    const int grid = (int)std::min<int64_t>((nw + 255) / 256, (int64_t)ctx->num_sms * 8);
    iq_snapshot_kernel<<<grid, 256, 0, ctx->stream>>>((const uint8_t *)d_in, nw, (int16_t *)d_i, (int16_t *)d_q);
    MKID_CHECK_LAUNCH(ctx);
    if ((rc = mkid_stage_out_finish(ctx, I, nw * 4, d_i))) return rc;
    return mkid_stage_out_finish(ctx, Q, nw * 4, d_q);
}

extern "C" int mkid_phase_deg_from_iq(mkid_ctx *ctx, const int16_t *I, const int16_t *Q, int64_t n, double Ic, double Qc,
                                      double *deg) {
    if (!ctx) return MKID_EINVAL;
    MKID_REQUIRE(ctx, I && Q && deg && n >= 0, "phase_deg_from_iq: NULL argument");
    if (n == 0) return MKID_OK;
    MKID_CUDA(ctx, cudaSetDevice(ctx->device));
    const void *d_i, *d_q; void *d_o; int rc;
    if ((rc = mkid_stage_in(ctx, I, n * 2, SCR_IN, &d_i))) return rc;
    if ((rc = mkid_stage_in(ctx, Q, n * 2, SCR_IN1, &d_q))) return rc;
    if ((rc = mkid_stage_out(ctx, deg, n * 8, SCR_OUT0, false, &d_o))) return rc;
    const int grid = (int)std::min<int64_t>((n + 255) / 256, (int64_t)ctx->num_sms * 8);
    phase_deg_kernel<<<grid, 256, 0, ctx->stream>>>((const int16_t *)d_i, (const int16_t *)d_q, n, Ic, Qc, (double *)d_o);
    MKID_CHECK_LAUNCH(ctx);
    return mkid_stage_out_finish(ctx, deg, n * 8, d_o);
}

extern "C" int mkid_soft_trigger(mkid_ctx *ctx, const double *phase, int32_t n_streams, int64_t n, const mkid_trigger_cfg *cfg,
                                 int32_t *hits, int32_t max_hits, int32_t *n_hits) {
    if (!ctx) return MKID_EINVAL;
    MKID_REQUIRE(ctx, phase && cfg && hits && n_hits && n_streams > 0 && n > 0 && max_hits > 0, "soft_trigger: bad argument");
    MKID_REQUIRE(ctx, (cfg->mode == 0 || cfg->mode == 1) && cfg->mean_len >= 1 && cfg->holdoff >= 1 && cfg->start >= 0 &&
                          cfg->tail >= 0, "soft_trigger: bad configuration");
    MKID_REQUIRE(ctx, cfg->mode != 0 || cfg->start >= cfg->mean_len, "soft_trigger: rolling mean needs start >= mean_len");
    MKID_REQUIRE(ctx, n < ((int64_t)1 << 31), "soft_trigger: at most 2^31 samples per stream");
    MKID_CUDA(ctx, cudaSetDevice(ctx->device));
    int rc;
    const void *d_x; void *d_hits, *d_nh;
    if ((rc = mkid_stage_in(ctx, phase, (size_t)n_streams * n * 8, SCR_IN, &d_x))) return rc;
    if ((rc = mkid_stage_out(ctx, hits, (size_t)n_streams * max_hits * 4, SCR_OUT0, false, &d_hits))) return rc;
    if ((rc = mkid_stage_out(ctx, n_hits, (size_t)n_streams * 4, SCR_OUT1, false, &d_nh))) return rc;
    TrigParams p;
    p.x = (const double *)d_x; p.n = n; p.n_streams = n_streams; p.mode = cfg->mode; p.M = cfg->mean_len;
    p.wrap = cfg->mode == 1 && cfg->wrap_negative; p.seq = cfg->sum_order != 0; p.thr = cfg->threshold;
    p.n_mask_words = (n + 31) / 32;
    void *d_mask, *d_means = nullptr;
    if ((rc = mkid_scratch(ctx, SCR_AUX2, (size_t)n_streams * p.n_mask_words * 4, &d_mask))) return rc;
    p.mask = (uint32_t *)d_mask;
    if (cfg->mode == 1) {
        const int64_t n_means = n / cfg->mean_len;
        if ((rc = mkid_scratch(ctx, SCR_AUX3, (size_t)n_streams * std::max<int64_t>(n_means, 1) * 8, &d_means))) return rc;
        p.means = (double *)d_means;
        if (n_means > 0) {
            trig_block_means_kernel<<<dim3((unsigned)((n_means + 127) / 128), n_streams), 128, 0, ctx->stream>>>(p);
            MKID_CHECK_LAUNCH(ctx);
        }
    } else p.means = nullptr;
    trig_candidates_kernel<<<dim3((unsigned)((n + 255) / 256), n_streams), 256, 0, ctx->stream>>>(p);
    MKID_CHECK_LAUNCH(ctx);
    trig_greedy_kernel<<<(n_streams + 3) / 4, 128, 0, ctx->stream>>>(p.mask, p.n_mask_words, n, n_streams, cfg->start,
                                                                   cfg->holdoff, cfg->tail, (int32_t *)d_hits, max_hits,
                                                                   (int32_t *)d_nh);
    MKID_CHECK_LAUNCH(ctx);
    if ((rc = mkid_stage_out_finish(ctx, hits, (size_t)n_streams * max_hits * 4, d_hits))) return rc;
    return mkid_stage_out_finish(ctx, n_hits, (size_t)n_streams * 4, d_nh);
}

extern "C" int mkid_thresholds_from_phase(mkid_ctx *ctx, const int16_t *phase, int32_t n_boards, int64_t board_stride,
                                          int64_t row_stride, int32_t n_ch, int64_t n_samples, double nsigma,
                                          int32_t *thr_raw, double *med, double *p5) {
    if (!ctx) return MKID_EINVAL;
    MKID_REQUIRE(ctx, phase && thr_raw && n_boards > 0 && n_ch > 0 && n_samples > 0 && row_stride >= n_ch,
                 "thresholds_from_phase: bad argument");
    MKID_REQUIRE(ctx, n_samples < (1 << 24), "thresholds_from_phase: at most 2^24 samples per channel (float32 counts)");
    MKID_REQUIRE(ctx, mkid_is_device_ptr(phase), "thresholds_from_phase: phase must be device memory (strided view)");
    MKID_CUDA(ctx, cudaSetDevice(ctx->device));
    int rc;
    void *d_thr, *d_med = nullptr, *d_p5 = nullptr;
    const size_t cnt = (size_t)n_boards * n_ch;
    if ((rc = mkid_stage_out(ctx, thr_raw, cnt * 4, SCR_OUT0, false, &d_thr))) return rc;
    if (med && (rc = mkid_stage_out(ctx, med, cnt * 8, SCR_OUT1, false, &d_med))) return rc;
    if (p5 && (rc = mkid_stage_out(ctx, p5, cnt * 8, SCR_OUT2, false, &d_p5))) return rc;
    thresholds_kernel<<<dim3(n_ch, n_boards), 256, 0, ctx->stream>>>(phase, board_stride, row_stride, n_ch, n_samples, nsigma,
                                                                      -25736, (int32_t *)d_thr, (double *)d_med, (double *)d_p5);
    MKID_CHECK_LAUNCH(ctx);
    if ((rc = mkid_stage_out_finish(ctx, thr_raw, cnt * 4, d_thr))) return rc;
    if (med && (rc = mkid_stage_out_finish(ctx, med, cnt * 8, d_med))) return rc;
    if (p5 && (rc = mkid_stage_out_finish(ctx, p5, cnt * 8, d_p5))) return rc;
    return MKID_OK;
}
