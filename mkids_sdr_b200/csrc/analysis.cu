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

// ---------------------------------------------------------------- longsnapshot noise spectrum (ROACH_Pulses.py:521-537)
// noiseFFT[k] = mean over the nFFTAverages segments of 20*log10(|fft(segment)[k]| / norm / 1e-6).  The segment
// length nLongsnapSamples/100 = 10485 is not a power of two: direct DFT with an exact twiddle table
// (k*n mod N in integers, sincospi), float64; one CTA per (output bin block, stream).
namespace {
__global__ void dft_twiddle_kernel(int N, double2 *tw) {
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= N) return;
    double s, c;
    sincospi(2.0 * (double)m / (double)N, &s, &c);
    tw[m] = make_double2(c, -s);                                   // e^{-2 pi i m / N}
}
__global__ void __launch_bounds__(128) noise_spectrum_kernel(const double *__restrict__ x, int64_t stream_stride, int N,
                                                             int n_avg, double norm, const double2 *__restrict__ tw,
                                                             double *out) {
    extern __shared__ double s_x[];                                // one segment
    const int s = blockIdx.y;
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const double *xs = x + (size_t)s * stream_stride;
    double acc_db = 0.0;
    for (int a = 0; a < n_avg; ++a) {
        __syncthreads();
        for (int i = threadIdx.x; i < N; i += blockDim.x) s_x[i] = xs[(size_t)a * N + i];
        __syncthreads();
        if (k < N) {
            double re = 0.0, im = 0.0;
            int m = 0;                                             // k*n mod N
            for (int n = 0; n < N; ++n) {
                const double2 w = tw[m];
                re = fma(s_x[n], w.x, re);
                im = fma(s_x[n], w.y, im);
                m += k; if (m >= N) m -= N;
            }
            const double mag = sqrt(re * re + im * im);
            acc_db += 20.0 * log10(mag / norm / 1e-6);
        }
    }
    if (k < N) out[(size_t)s * N + k] = acc_db / (double)n_avg;
}
}  // namespace

extern "C" int mkid_noise_spectrum(mkid_ctx *ctx, const double *phase_deg, int32_t n_streams, int64_t n_samples,
                                   int32_t n_averages, double norm, double *noise_db) {
    if (!ctx) return MKID_EINVAL;
    MKID_REQUIRE(ctx, phase_deg && noise_db && n_streams > 0 && n_averages > 0 && n_samples >= n_averages && norm > 0.0,
                 "noise_spectrum: bad argument");
    const int N = (int)(n_samples / n_averages);                   // nSamplesPerFFT (integer division, :522)
    MKID_REQUIRE(ctx, N >= 1 && (size_t)N * 8 <= 200 * 1024, "noise_spectrum: segment must fit shared memory (<= 25600 samples)");
    MKID_CUDA(ctx, cudaSetDevice(ctx->device));
    int rc;
    const void *d_x; void *d_o, *d_tw;
    if ((rc = mkid_stage_in(ctx, phase_deg, (size_t)n_streams * n_samples * 8, SCR_IN, &d_x))) return rc;
    if ((rc = mkid_stage_out(ctx, noise_db, (size_t)n_streams * N * 8, SCR_OUT0, false, &d_o))) return rc;
    if ((rc = mkid_scratch(ctx, SCR_AUX2, (size_t)N * 16, &d_tw))) return rc;
    dft_twiddle_kernel<<<(N + 255) / 256, 256, 0, ctx->stream>>>(N, (double2 *)d_tw);
    MKID_CHECK_LAUNCH(ctx);
    const size_t smem = (size_t)N * 8;
    MKID_CUDA(ctx, cudaFuncSetAttribute(noise_spectrum_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    noise_spectrum_kernel<<<dim3((N + 127) / 128, n_streams), 128, smem, ctx->stream>>>((const double *)d_x, n_samples, N, n_averages,
                                                                                      norm, (const double2 *)d_tw, (double *)d_o);
    MKID_CHECK_LAUNCH(ctx);
    return mkid_stage_out_finish(ctx, noise_db, (size_t)n_streams * N * 8, d_o);
}

// ---------------------------------------------------------------- image_Worker spectra products (ArconsDashboard.py:1282-1504)
// darray u32 [n_pix][10] (the per-pixel 10-bin spectrum K6 accumulates).  Per bin: numpy.median over the pixels
// (bitonic sort in shared memory), optional sky subtraction x - int(median); pc[p] = sum of the 10 bins;
// me[p] = (C0*E0 + ... + C9*E9) / pc[p] in the reference's left-to-right float64 order, h*c/me for wavelength bins.
namespace {
__global__ void __launch_bounds__(1024) spectra_median_kernel(const uint32_t *__restrict__ darray, int n_pix, int n_pow2,
                                                              double *medians) {
    extern __shared__ uint32_t s_v[];
    const int bin = blockIdx.x, tid = threadIdx.x;
    for (int i = tid; i < n_pow2; i += 1024) s_v[i] = i < n_pix ? darray[(size_t)i * 10 + bin] : 0xFFFFFFFFu;
    __syncthreads();
    for (int k = 2; k <= n_pow2; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < n_pow2; i += 1024) {
                const int l = i ^ j;
                if (l > i) {
                    const uint32_t a = s_v[i], b = s_v[l];
                    const bool up = (i & k) == 0;
                    if ((a > b) == up) { s_v[i] = b; s_v[l] = a; }
                }
            }
            __syncthreads();
        }
    if (tid == 0) {
        // numpy.median: mean of the two middle elements for even n
        const double m = (n_pix & 1) ? (double)s_v[n_pix / 2] : ((double)s_v[n_pix / 2 - 1] + (double)s_v[n_pix / 2]) / 2.0;
        medians[bin] = m;
    }
}
__global__ void spectra_products_kernel(const uint32_t *__restrict__ darray, int n_pix, const double *__restrict__ medians,
                                        int sky, const double *__restrict__ E, double hc, int wavelength, long long *pc,
                                        double *me) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_pix) return;
    long long sum = 0;
    double num = 0.0;
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        long long cv = (long long)darray[(size_t)p * 10 + i];
        if (sky) cv -= (long long)medians[i];                       // x - int(meds[i])
        sum += cv;
        const double term = __dmul_rn((double)cv, E[i]);
        num = i == 0 ? term : __dadd_rn(num, term);
    }
    pc[p] = sum;
    double m = __ddiv_rn(num, (double)sum);                          // inf / nan for an empty pixel, as numpy
    if (wavelength) m = __ddiv_rn(hc, m);
    me[p] = m;
}
}  // namespace

extern "C" int mkid_spectra_products(mkid_ctx *ctx, const uint32_t *darray, int32_t n_pix, const double *bin_centres,
                                     double hc, int32_t wavelength, int32_t sky_subtraction, double *medians,
                                     int64_t *pc, double *me) {
    if (!ctx) return MKID_EINVAL;
    MKID_REQUIRE(ctx, darray && bin_centres && medians && pc && me && n_pix > 0, "spectra_products: NULL argument");
    int n_pow2 = 1;
    while (n_pow2 < n_pix) n_pow2 <<= 1;
    MKID_REQUIRE(ctx, (size_t)n_pow2 * 4 <= 200 * 1024, "spectra_products: at most 51200 pixels");
    MKID_CUDA(ctx, cudaSetDevice(ctx->device));
    int rc;
    const void *d_in, *d_E; void *d_med, *d_pc, *d_me;
    if ((rc = mkid_stage_in(ctx, darray, (size_t)n_pix * 40, SCR_IN, &d_in))) return rc;
    if ((rc = mkid_stage_in(ctx, bin_centres, 80, SCR_IN1, &d_E))) return rc;
    if ((rc = mkid_stage_out(ctx, medians, 80, SCR_OUT0, false, &d_med))) return rc;
    if ((rc = mkid_stage_out(ctx, pc, (size_t)n_pix * 8, SCR_OUT1, false, &d_pc))) return rc;
    if ((rc = mkid_stage_out(ctx, me, (size_t)n_pix * 8, SCR_OUT2, false, &d_me))) return rc;
    const size_t smem = (size_t)n_pow2 * 4;
    MKID_CUDA(ctx, cudaFuncSetAttribute(spectra_median_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    spectra_median_kernel<<<10, 1024, smem, ctx->stream>>>((const uint32_t *)d_in, n_pix, n_pow2, (double *)d_med);
    MKID_CHECK_LAUNCH(ctx);
    spectra_products_kernel<<<(n_pix + 255) / 256, 256, 0, ctx->stream>>>((const uint32_t *)d_in, n_pix, (const double *)d_med,
                                                                         sky_subtraction, (const double *)d_E, hc, wavelength,
                                                                         (long long *)d_pc, (double *)d_me);
    MKID_CHECK_LAUNCH(ctx);
    if ((rc = mkid_stage_out_finish(ctx, medians, 80, d_med))) return rc;
    if ((rc = mkid_stage_out_finish(ctx, pc, (size_t)n_pix * 8, d_pc))) return rc;
    return mkid_stage_out_finish(ctx, me, (size_t)n_pix * 8, d_me);
}

// ---------------------------------------------------------------- dashboard image (ArconsDashboard.py:633-723 make_image)
// Per-second quick-look images are gathers of the capped per-(second,pixel) counts through the beammap
// (PacketMaster.c:1029-1045, uint16).  The dashboard sums the images of the seconds [t_i, t_f) (or shows second t_f
// alone when t_f == 0 or t_f == t_i, :673-677), subtracts skyrate*(t_f - t_i) (:679-680), flips the rows back (:683)
// and applies the flat field (:688-689).  All of it in one gather kernel, float64 like the NumPy code.
namespace {
__global__ void dashboard_image_kernel(const uint32_t *__restrict__ counts, int n_pix, const int32_t *__restrict__ pixel_adr,
                                       int rows, int cols, int t_i, int t_f, int max_events, const double *__restrict__ skyrate,
                                       const double *__restrict__ flat, double *image, double *image_counts) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;          // index into the (rows x cols) text image of a second
    if (i >= rows * cols) return;
    const int r = i / cols, c = i - r * cols;
    const int adr = pixel_adr[i];
    const uint32_t cap = (uint32_t)(max_events - 1);
    double v;
    if (t_f == 0 || t_f == t_i) {
        v = (double)(uint16_t)min(counts[(size_t)t_f * n_pix + adr], cap);
    } else {
        v = 0.0;
        for (int s = t_i; s < t_f; ++s) v += (double)(uint16_t)min(counts[(size_t)s * n_pix + adr], cap);   // sum(self.counts[ti:tf])
    }
    if (skyrate) v = __dsub_rn(v, __dmul_rn(skyrate[i], (double)(t_f - t_i)));
    image_counts[i] = v;
    // photon_count = flipud(reshape(image_counts, rawshape)): row r of the displayed frame is row rows-1-r of the text image
    double pc = v;
    const int o = (rows - 1 - r) * cols + c;
    if (flat) pc = __dmul_rn(pc, flat[o]);
    image[o] = pc;
}
}  // namespace

extern "C" int mkid_dashboard_image(mkid_ctx *ctx, const uint32_t *counts_raw, int32_t n_pix, const int32_t *pixel_adr,
                                    int32_t rows, int32_t cols, int32_t t_i, int32_t t_f, int32_t max_events,
                                    const double *skyrate, const double *flat, double *image, double *image_counts) {
    if (!ctx) return MKID_EINVAL;
    MKID_REQUIRE(ctx, counts_raw && pixel_adr && image && image_counts && rows > 0 && cols > 0 && n_pix > 0 && t_i >= 0 &&
                          t_f >= t_i && max_events >= 2, "dashboard_image: bad argument");
    MKID_REQUIRE(ctx, mkid_is_device_ptr(counts_raw), "dashboard_image: counts_raw must be device memory ([exptime][n_pix])");
    MKID_CUDA(ctx, cudaSetDevice(ctx->device));
    const int n = rows * cols;
    int rc;
    const void *d_adr, *d_sky = nullptr, *d_flat = nullptr; void *d_img, *d_ic;
    if ((rc = mkid_stage_in(ctx, pixel_adr, (size_t)n * 4, SCR_IN, &d_adr))) return rc;
    if (skyrate && (rc = mkid_stage_in(ctx, skyrate, (size_t)n * 8, SCR_IN1, &d_sky))) return rc;
    if (flat && (rc = mkid_stage_in(ctx, flat, (size_t)n * 8, SCR_IN2, &d_flat))) return rc;
    if ((rc = mkid_stage_out(ctx, image, (size_t)n * 8, SCR_OUT0, false, &d_img))) return rc;
    if ((rc = mkid_stage_out(ctx, image_counts, (size_t)n * 8, SCR_OUT1, false, &d_ic))) return rc;
    dashboard_image_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(counts_raw, n_pix, (const int32_t *)d_adr, rows, cols, t_i, t_f,
                                                                  max_events, (const double *)d_sky, (const double *)d_flat,
                                                                  (double *)d_img, (double *)d_ic);
    MKID_CHECK_LAUNCH(ctx);
    if ((rc = mkid_stage_out_finish(ctx, image, (size_t)n * 8, d_img))) return rc;
    return mkid_stage_out_finish(ctx, image_counts, (size_t)n * 8, d_ic);
}
