// Matched-filter template builder: the per-pulse arithmetic of MakeTemplate,
// DataReadout/ReadoutControls/lib/pulses.py:239-427, batched over all pulses of a resonator.
//
//   mkid_tpl_median      numpy.median of a sub-block of the float32 pulse table (:273-274)
//   mkid_tpl_prepare     per pulse: median de-trend in place (:283-284), arctan2 (:287), unwrap -> degrees (:291),
//                        straight-line baseline fit on samples [0,900) + [1800,2000) (:294-295), statistics for the
//                        rejection tests (:298-313), peak search
//   mkid_tpl_convpeak    arg-max of numpy.convolve(tP[900:1500], P3) (:355-357)
//   mkid_tpl_accumulate  template sum  tP += roll(P3, shift)/max(P3)  in pulse order (:319-320, :371-372)
//   mkid_tpl_noise       noise += |fft(deg2rad(P4[50:850]))|^2 in pulse order (:376)
//
// The accept / reject decisions on the per-pulse scalars (a handful of comparisons per pulse) stay on the host, as
// in the reference (mkids_sdr_b200/template.py).  float32 where NumPy computes in float32 (the table, arctan2,
// unwrap, rad2deg), float64 from the baseline fit on.  Sums that NumPy evaluates left to right over the pulses are
// evaluated left to right here; the per-pulse reductions (fit, convolution, DFT) are parallel: parity is by
// tolerance (tests/test_template_gpu.py), not bit-exact, because NumPy's float32 arctan2 is libm's.
#include <math.h>

#include "common.cuh"

namespace {

constexpr int TPL_N = 2000;           // samples per pulse (:254)
constexpr int TPL_K = 600;            // correlation kernel tP[900:1500]
constexpr int TPL_NOISE = 800;        // P4[50:850]

__device__ __forceinline__ uint32_t f2key(float f) { const uint32_t u = __float_as_uint(f); return (u & 0x80000000u) ? ~u : (u | 0x80000000u); }
__device__ __forceinline__ float key2f(uint32_t k) { return __uint_as_float((k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k); }

// k-th smallest (0-based) of n floats addressed by get(i): 4 passes of an 8-bit radix select, one CTA
template <typename GET>
__device__ float select_kth(GET get, int n, int k, unsigned *s_hist, unsigned *s_sel) {
    uint32_t prefix = 0, mask = 0;
    for (int pass = 3; pass >= 0; --pass) {
        for (int i = threadIdx.x; i < 256; i += blockDim.x) s_hist[i] = 0;
        __syncthreads();
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const uint32_t key = f2key(get(i));
            if ((key & mask) == prefix) atomicAdd(&s_hist[(key >> (8 * pass)) & 255u], 1u);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            int acc = 0, b = 0;
            for (; b < 256; ++b) { if (acc + (int)s_hist[b] > k) break; acc += s_hist[b]; }
            s_sel[0] = (unsigned)b; s_sel[1] = (unsigned)acc;
        }
        __syncthreads();
        prefix |= s_sel[0] << (8 * pass);
        mask |= 255u << (8 * pass);
        k -= (int)s_sel[1];
        __syncthreads();
    }
    return key2f(prefix);
}

__global__ void __launch_bounds__(1024) tpl_median_kernel(const float *a, int rows, int cols, int64_t row_stride, float *out) {
    __shared__ unsigned s_hist[256], s_sel[2];
    const int n = rows * cols;
    auto get = [&](int i) { return a[(size_t)(i / cols) * row_stride + (i % cols)]; };
    float m;
    if (n & 1) m = select_kth(get, n, n / 2, s_hist, s_sel);
    else {
        const float lo = select_kth(get, n, n / 2 - 1, s_hist, s_sel), hi = select_kth(get, n, n / 2, s_hist, s_sel);
        m = __fdiv_rn(__fadd_rn(lo, hi), 2.0f);          // numpy.median: mean of the two middle elements, float32
    }
    if (threadIdx.x == 0) *out = m;
}

__device__ __forceinline__ float np_modf32(float a, float b) {         // numpy.mod for floats (b > 0)
    float m = fmodf(a, b);
    if (m != 0.f) { if (m < 0.f) m += b; } else m = 0.f;
    return m;
}

struct PulseStats { double mean_first, mean_last, std_first, peak, max_all; int32_t ploc, pad; };

__global__ void __launch_bounds__(256) tpl_prepare_kernel(float *I_all, float *Q_all, float I1m, float Q1m, double *P3_all,
                                                          PulseStats *stats) {
    __shared__ float sI[TPL_N], sQ[TPL_N];
    __shared__ double sP[TPL_N];
    __shared__ unsigned s_hist[256], s_sel[2];
    __shared__ double s_red[5][8];
    const int j = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float *I = I_all + (size_t)j * TPL_N, *Q = Q_all + (size_t)j * TPL_N;
    for (int t = tid; t < TPL_N; t += 256) { sI[t] = I[t]; sQ[t] = Q[t]; }
    __syncthreads();
    // reference all pulses to the first 100 pulses (1/f removal): I += I1m - median(I[1:900]), in place
    const float mI = select_kth([&](int i) { return sI[1 + i]; }, 899, 449, s_hist, s_sel);
    const float mQ = select_kth([&](int i) { return sQ[1 + i]; }, 899, 449, s_hist, s_sel);
    const float dI = __fsub_rn(I1m, mI), dQ = __fsub_rn(Q1m, mQ);
    for (int t = tid; t < TPL_N; t += 256) {
        const float vi = __fadd_rn(sI[t], dI), vq = __fadd_rn(sQ[t], dQ);
        I[t] = vi; Q[t] = vq;
        sI[t] = atan2f(vq, vi);                           // P1 (float32); xc = yc = 0
    }
    __syncthreads();
    // numpy.unwrap in float32, then rad2deg
    if (tid == 0) {
        const float pi_f = 3.14159265358979323846f, per = 6.28318530717958647692f;
        float cs = 0.f;
        float prev = sI[0];
        sQ[0] = __fmul_rn(prev, 57.29577951308232f);
        for (int t = 1; t < TPL_N; ++t) {
            const float cur = sI[t];
            const float dd = __fsub_rn(cur, prev);
            float ddmod = __fadd_rn(np_modf32(__fadd_rn(dd, pi_f), per), -pi_f);
            if (ddmod == -pi_f && dd > 0.f) ddmod = pi_f;
            float corr = __fsub_rn(ddmod, dd);
            if (fabsf(dd) < pi_f) corr = 0.f;
            cs = __fadd_rn(cs, corr);
            sQ[t] = __fmul_rn(__fadd_rn(cur, cs), 57.29577951308232f);       // P2 = rad2deg(unwrap(P1))
            prev = cur;
        }
    }
    __syncthreads();
    // straight-line fit (numpy.polyfit degree 1) over x = 2t, t in [0,900) + [1800,2000): 1100 points
    double sy = 0.0, sxy = 0.0;
    const double xbar = 1568900.0 / 1100.0;              // mean of the 1100 abscissae 2t: 2*(404550 + 379900)/1100
    for (int t = tid; t < TPL_N; t += 256)
        if (t < 900 || t >= 1800) { const double y = (double)sQ[t]; sy += y; sxy += (2.0 * t - xbar) * y; }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) { sy += __shfl_xor_sync(0xffffffffu, sy, d); sxy += __shfl_xor_sync(0xffffffffu, sxy, d); }
    if (lane == 0) { s_red[0][warp] = sy; s_red[1][warp] = sxy; }
    __syncthreads();
    double SY = 0.0, SXY = 0.0;
    for (int w = 0; w < 8; ++w) { SY += s_red[0][w]; SXY += s_red[1][w]; }
    double sxx = 0.0;                                     // sum (x - xbar)^2 (the same for every pulse)
    for (int t = 0; t < TPL_N; ++t) if (t < 900 || t >= 1800) sxx += (2.0 * t - xbar) * (2.0 * t - xbar);
    const double slope = SXY / sxx, icpt = SY / 1100.0 - slope * xbar;
    for (int t = tid; t < TPL_N; t += 256) {
        const double v = (double)sQ[t] - (slope * (2.0 * t) + icpt);          // P3 = P2 - fit(idx)
        sP[t] = v;
        P3_all[(size_t)j * TPL_N + t] = v;
    }
    __syncthreads();
    // statistics: mean / std of P3[:100], mean of P3[1900:], max of P3[980:1050] and its first position, max(P3)
    double a0 = 0.0, a1 = 0.0, pk = -1e300, mx = -1e300;
    for (int t = tid; t < TPL_N; t += 256) {
        const double v = sP[t];
        if (t < 100) a0 += v;
        if (t >= 1900) a1 += v;
        if (t >= 980 && t < 1050) pk = fmax(pk, v);
        mx = fmax(mx, v);
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        a0 += __shfl_xor_sync(0xffffffffu, a0, d); a1 += __shfl_xor_sync(0xffffffffu, a1, d);
        pk = fmax(pk, __shfl_xor_sync(0xffffffffu, pk, d)); mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, d));
    }
    if (lane == 0) { s_red[0][warp] = a0; s_red[1][warp] = a1; s_red[2][warp] = pk; s_red[3][warp] = mx; }
    __syncthreads();
    double A0 = 0.0, A1 = 0.0, PK = -1e300, MX = -1e300;
    for (int w = 0; w < 8; ++w) { A0 += s_red[0][w]; A1 += s_red[1][w]; PK = fmax(PK, s_red[2][w]); MX = fmax(MX, s_red[3][w]); }
    const double m0 = A0 / 100.0;
    double var = 0.0;
    for (int t = tid; t < 100; t += 256) var += (sP[t] - m0) * (sP[t] - m0);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) var += __shfl_xor_sync(0xffffffffu, var, d);
    __syncthreads();
    if (lane == 0) s_red[4][warp] = var;
    // first index with P3 == peak (np.where(P3 == peak)[0])
    __shared__ int s_ploc;
    if (tid == 0) s_ploc = TPL_N;
    __syncthreads();
    for (int t = tid; t < TPL_N; t += 256) if (sP[t] == PK) atomicMin(&s_ploc, t);
    __syncthreads();
    if (tid == 0) {
        double V = 0.0;
        for (int w = 0; w < 8; ++w) V += s_red[4][w];
        PulseStats st;
        st.mean_first = m0; st.mean_last = A1 / 100.0; st.std_first = sqrt(V / 100.0); st.peak = PK; st.max_all = MX;
        st.ploc = s_ploc; st.pad = 0;
        stats[j] = st;
    }
}

// conv = numpy.convolve(kern[600], P3[2000]) (full, 2599 points): first arg-max and P3[1000 + argmax - 1160]
__global__ void __launch_bounds__(256) tpl_convpeak_kernel(const double *__restrict__ P3_all, const double *__restrict__ kern,
                                                           int32_t *argmax, double *p3_at) {
    __shared__ double sP[TPL_N], sK[TPL_K];
    __shared__ double s_best[8];
    __shared__ int s_idx[8];
    const int j = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int t = tid; t < TPL_N; t += 256) sP[t] = P3_all[(size_t)j * TPL_N + t];
    for (int t = tid; t < TPL_K; t += 256) sK[t] = kern[t];
    __syncthreads();
    double best = -1e300; int bi = 0x7fffffff;
    for (int k = tid; k < TPL_N + TPL_K - 1; k += 256) {
        const int m0 = max(0, k - (TPL_N - 1)), m1 = min(TPL_K - 1, k);
        double acc = 0.0;
        for (int m = m0; m <= m1; ++m) acc = fma(sK[m], sP[k - m], acc);
        if (acc > best) { best = acc; bi = k; }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        const double ob = __shfl_xor_sync(0xffffffffu, best, d); const int oi = __shfl_xor_sync(0xffffffffu, bi, d);
        if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
    }
    if (lane == 0) { s_best[warp] = best; s_idx[warp] = bi; }
    __syncthreads();
    if (tid == 0) {
        for (int w = 1; w < 8; ++w) if (s_best[w] > best || (s_best[w] == best && s_idx[w] < bi)) { best = s_best[w]; bi = s_idx[w]; }
        argmax[j] = bi;
        const int at = 1000 + bi - 1160;
        p3_at[j] = (at >= 0 && at < TPL_N) ? sP[at] : nan("");
    }
}

// out[t] (+)= sum over the listed pulses, in list order, of P3[pulse][(t - shift) mod 2000] / max_all[pulse]
__global__ void tpl_accumulate_kernel(const double *__restrict__ P3_all, const int32_t *__restrict__ pulse, const int32_t *__restrict__ shift,
                                      const double *__restrict__ norm, int n_list, double *out) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= TPL_N) return;
    double acc = out[t];
    for (int i = 0; i < n_list; ++i) {
        int src = (t - shift[i]) % TPL_N;
        if (src < 0) src += TPL_N;
        acc = __dadd_rn(acc, __ddiv_rn(P3_all[(size_t)pulse[i] * TPL_N + src], norm[i]));       // tP += P4/np.max(P4)
    }
    out[t] = acc;
}

// psd[i][k] = |fft(deg2rad(P4[50:850]))[k]|^2 of listed pulse i (P4 = roll(P3, shift)), direct 800-point DFT
__global__ void __launch_bounds__(256) tpl_psd_kernel(const double *__restrict__ P3_all, const int32_t *__restrict__ pulse,
                                                      const int32_t *__restrict__ shift, double *psd) {
    __shared__ double sx[TPL_NOISE];
    __shared__ double2 stw[TPL_NOISE];
    const int i = blockIdx.x, tid = threadIdx.x;
    for (int n = tid; n < TPL_NOISE; n += 256) {
        int src = (50 + n - shift[i]) % TPL_N;
        if (src < 0) src += TPL_N;
        sx[n] = __dmul_rn(P3_all[(size_t)pulse[i] * TPL_N + src], 0.017453292519943295);       // deg2rad
        double s, c;
        sincospi(2.0 * (double)n / (double)TPL_NOISE, &s, &c);
        stw[n] = make_double2(c, -s);
    }
    __syncthreads();
    for (int k = tid; k < TPL_NOISE; k += 256) {
        double re = 0.0, im = 0.0;
        int m = 0;
        for (int n = 0; n < TPL_NOISE; ++n) {
            re = fma(sx[n], stw[m].x, re); im = fma(sx[n], stw[m].y, im);
            m += k; if (m >= TPL_NOISE) m -= TPL_NOISE;
        }
        psd[(size_t)i * TPL_NOISE + k] = re * re + im * im;
    }
}
__global__ void tpl_noise_sum_kernel(const double *__restrict__ psd, int n_list, double *noise) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= TPL_NOISE) return;
    double acc = noise[k];
    for (int i = 0; i < n_list; ++i) acc = __dadd_rn(acc, psd[(size_t)i * TPL_NOISE + k]);
    noise[k] = acc;
}

}  // namespace

extern "C" int mkid_tpl_median(mkid_ctx *ctx, const float *table, int32_t rows, int32_t cols, int64_t row_stride, float *out) {
    if (!ctx) return MKID_EINVAL;
    MKID_REQUIRE(ctx, table && out && rows > 0 && cols > 0 && row_stride >= cols && mkid_is_device_ptr(table),
                 "tpl_median: table must be a device pointer");
    MKID_CUDA(ctx, cudaSetDevice(ctx->device));
    void *d_o; int rc;
    if ((rc = mkid_stage_out(ctx, out, 4, SCR_OUT0, false, &d_o))) return rc;
    tpl_median_kernel<<<1, 1024, 0, ctx->stream>>>(table, rows, cols, row_stride, (float *)d_o);
    MKID_CHECK_LAUNCH(ctx);
    return mkid_stage_out_finish(ctx, out, 4, d_o);
}

extern "C" int mkid_tpl_prepare(mkid_ctx *ctx, float *I, float *Q, int32_t n_pulses, float I1m, float Q1m, double *P3,
                                double *stats) {
    if (!ctx) return MKID_EINVAL;
    MKID_REQUIRE(ctx, I && Q && P3 && stats && n_pulses > 0, "tpl_prepare: NULL argument");
    MKID_REQUIRE(ctx, mkid_is_device_ptr(I) && mkid_is_device_ptr(Q) && mkid_is_device_ptr(P3), "tpl_prepare: I, Q, P3 must be device memory");
    MKID_CUDA(ctx, cudaSetDevice(ctx->device));
    void *d_s; int rc;
    static_assert(sizeof(PulseStats) == 48, "PulseStats layout is part of the ABI (6 doubles per pulse)");
    if ((rc = mkid_stage_out(ctx, stats, (size_t)n_pulses * sizeof(PulseStats), SCR_OUT0, false, &d_s))) return rc;
    tpl_prepare_kernel<<<n_pulses, 256, 0, ctx->stream>>>(I, Q, I1m, Q1m, P3, (PulseStats *)d_s);
    MKID_CHECK_LAUNCH(ctx);
    return mkid_stage_out_finish(ctx, stats, (size_t)n_pulses * sizeof(PulseStats), d_s);
}

extern "C" int mkid_tpl_convpeak(mkid_ctx *ctx, const double *P3, int32_t n_pulses, const double *kernel600, int32_t *argmax,
                                 double *p3_at) {
    if (!ctx) return MKID_EINVAL;
    MKID_REQUIRE(ctx, P3 && kernel600 && argmax && p3_at && n_pulses > 0 && mkid_is_device_ptr(P3), "tpl_convpeak: bad argument");
    MKID_CUDA(ctx, cudaSetDevice(ctx->device));
    const void *d_k; void *d_a, *d_p; int rc;
    if ((rc = mkid_stage_in(ctx, kernel600, TPL_K * 8, SCR_IN, &d_k))) return rc;
    if ((rc = mkid_stage_out(ctx, argmax, (size_t)n_pulses * 4, SCR_OUT0, false, &d_a))) return rc;
    if ((rc = mkid_stage_out(ctx, p3_at, (size_t)n_pulses * 8, SCR_OUT1, false, &d_p))) return rc;
    tpl_convpeak_kernel<<<n_pulses, 256, 0, ctx->stream>>>(P3, (const double *)d_k, (int32_t *)d_a, (double *)d_p);
    MKID_CHECK_LAUNCH(ctx);
    if ((rc = mkid_stage_out_finish(ctx, argmax, (size_t)n_pulses * 4, d_a))) return rc;
    return mkid_stage_out_finish(ctx, p3_at, (size_t)n_pulses * 8, d_p);
}

extern "C" int mkid_tpl_accumulate(mkid_ctx *ctx, const double *P3, const int32_t *pulse, const int32_t *shift, const double *norm,
                                   int32_t n_list, double *tmpl, double *noise) {
    if (!ctx) return MKID_EINVAL;
    MKID_REQUIRE(ctx, P3 && tmpl && mkid_is_device_ptr(P3) && n_list >= 0, "tpl_accumulate: bad argument");
    if (n_list == 0) return MKID_OK;
    MKID_REQUIRE(ctx, pulse && shift && norm, "tpl_accumulate: NULL list");
    MKID_CUDA(ctx, cudaSetDevice(ctx->device));
    const void *d_pu, *d_sh, *d_no; void *d_t, *d_n = nullptr; int rc;
    if ((rc = mkid_stage_in(ctx, pulse, (size_t)n_list * 4, SCR_IN, &d_pu))) return rc;
    if ((rc = mkid_stage_in(ctx, shift, (size_t)n_list * 4, SCR_IN1, &d_sh))) return rc;
    if ((rc = mkid_stage_in(ctx, norm, (size_t)n_list * 8, SCR_IN2, &d_no))) return rc;
    if ((rc = mkid_stage_out(ctx, tmpl, TPL_N * 8, SCR_OUT0, true, &d_t))) return rc;
    tpl_accumulate_kernel<<<(TPL_N + 127) / 128, 128, 0, ctx->stream>>>(P3, (const int32_t *)d_pu, (const int32_t *)d_sh,
                                                                       (const double *)d_no, n_list, (double *)d_t);
    MKID_CHECK_LAUNCH(ctx);
    if (noise) {
        void *d_psd;
        if ((rc = mkid_stage_out(ctx, noise, TPL_NOISE * 8, SCR_OUT1, true, &d_n))) return rc;
        if ((rc = mkid_scratch(ctx, SCR_AUX2, (size_t)n_list * TPL_NOISE * 8, &d_psd))) return rc;
        tpl_psd_kernel<<<n_list, 256, 0, ctx->stream>>>(P3, (const int32_t *)d_pu, (const int32_t *)d_sh, (double *)d_psd);
        MKID_CHECK_LAUNCH(ctx);
        tpl_noise_sum_kernel<<<(TPL_NOISE + 127) / 128, 128, 0, ctx->stream>>>((const double *)d_psd, n_list, (double *)d_n);
        MKID_CHECK_LAUNCH(ctx);
        if ((rc = mkid_stage_out_finish(ctx, noise, TPL_NOISE * 8, d_n))) return rc;
    }
    return mkid_stage_out_finish(ctx, tmpl, TPL_N * 8, d_t);
}
