// K1-K3: DAC frequency-comb LUT, DDS LUT and DRAM image synthesis.
//
// Replaces AppForm.freqCombLUT / define_DDS_LUT / write_LUTs of
// DataReadout/ChannelizerControls/ROACH_Setup.py:416-578 (multi-tone twin ROACH_Setup_DAC.py:396-558).
//
// K1  every comb frequency is a multiple of fs/N, so I + jQ = N * IDFT(X) with one spectral line
//     per tone.  Bulk: four-step fp64 IFFT, N = N2 x N1 (N1 a power of four <= 1024): for each
//     n2 < N2 a shared-memory Stockham IFFT of length N1 over the (sparse) lines (N1 = 1024, the
//     2^19-sample tables: radices 16 x 16 x 4 with the 16-point DFTs in registers; shorter tables:
//     radix 4), fused with the max search; then max candidates -> exact max -> scale -> tiled-
//     transpose quantise -> exact fix-up of the flagged samples.  The reference evaluates every sample as a sequential float64 sum of
//     a*cos(((2*pi*f)*t)/fs + phi); its int() truncation can flip on samples whose scaled value
//     is within the reference's own rounding error of an integer.  Those samples (and the
//     candidates for the max that sets the scale) are re-evaluated in the reference's operation
//     order with correctly rounded double-double sin/cos, so the int16 output is the reference's.
// K2  DDS tables are small (256 channels x N/256 samples): evaluated directly in reference order.
// K3  big-endian 8 x int16 DRAM image (ROACH_Setup.py:560-569).
#include <math.h>
#include <stdlib.h>

#include <algorithm>

#include "common.cuh"

namespace {

// ------------------------------------------------------------------ double-double arithmetic
struct dd { double hi, lo; };
__device__ __forceinline__ dd two_sum(double a, double b) {
    double s = __dadd_rn(a, b), bb = __dsub_rn(s, a);
    double e = __dadd_rn(__dsub_rn(a, __dsub_rn(s, bb)), __dsub_rn(b, bb));
    return {s, e};
}
__device__ __forceinline__ dd quick_two_sum(double a, double b) {
    double s = __dadd_rn(a, b);
    return {s, __dsub_rn(b, __dsub_rn(s, a))};
}
__device__ __forceinline__ dd two_prod(double a, double b) {
    double p = __dmul_rn(a, b);
    return {p, __fma_rn(a, b, -p)};
}
__device__ __forceinline__ dd dd_add(dd a, dd b) {
    dd s = two_sum(a.hi, b.hi);
    dd t = two_sum(a.lo, b.lo);
    s.lo = __dadd_rn(s.lo, t.hi);
    s = quick_two_sum(s.hi, s.lo);
    s.lo = __dadd_rn(s.lo, t.lo);
    return quick_two_sum(s.hi, s.lo);
}
__device__ __forceinline__ dd dd_mul(dd a, dd b) {
    dd p = two_prod(a.hi, b.hi);
    p.lo = __dadd_rn(p.lo, __dadd_rn(__dmul_rn(a.hi, b.lo), __dmul_rn(a.lo, b.hi)));
    return quick_two_sum(p.hi, p.lo);
}

__constant__ double c_S[13][2] = {   // (-1)^i / (2i+3)!  i = 0..12  (sin series after the leading r)
    {-0.16666666666666666, -9.25185853854297e-18},  {0.008333333333333333, 1.1564823173178714e-19},
    {-0.0001984126984126984, -1.7209558293420705e-22}, {2.7557319223985893e-06, -1.858393274046472e-22},
    {-2.505210838544172e-08, 1.448814070935912e-24}, {1.6059043836821613e-10, 1.2585294588752098e-26},
    {-7.647163731819816e-13, -7.03872877733453e-30}, {2.8114572543455206e-15, 1.6508842730861433e-31},
    {-8.22063524662433e-18, -2.2141894119604265e-34}, {1.9572941063391263e-20, -1.3643503830087908e-36},
    {-3.868170170630684e-23, 8.843177655482344e-40}, {6.446950284384474e-26, -1.9330404233703465e-42},
    {-9.183689863795546e-29, -1.4303150396787322e-45}};
__constant__ double c_C[13][2] = {   // (-1)^i / (2i)!  i = 1..13  (cos series after the leading 1)
    {-0.5, 0.0}, {0.041666666666666664, 2.3129646346357427e-18}, {-0.001388888888888889, 5.300543954373577e-20},
    {2.48015873015873e-05, 2.1511947866775882e-23}, {-2.755731922398589e-07, -2.3767714622250297e-23},
    {2.08767569878681e-09, -1.20734505911326e-25}, {-1.1470745597729725e-11, -2.0655512752830745e-28},
    {4.779477332387385e-14, 4.399205485834081e-31}, {-1.5619206968586225e-16, -1.1910679660273754e-32},
    {4.110317623312165e-19, 1.4412973378659527e-36}, {-8.896791392450574e-22, 7.911402614872376e-38},
    {1.6117375710961184e-24, -3.6846573564509766e-41}, {-2.4795962632247976e-27, 1.2953730964765229e-43}};

// sin and cos of a double, correctly rounded in all but astronomically rare cases (error < 2^-95
// relative before the final rounding), |x| < 2^30.
__device__ void sincos_cr(double x, double *s_out, double *c_out) {
    const double PIO2_1 = 1.5707963267948966, PIO2_2 = 6.123233995736766e-17, PIO2_3 = -1.4973849048591698e-33;
    const double k = rint(__dmul_rn(x, 0.6366197723675814));
    dd p = two_prod(k, PIO2_1);
    dd r = two_sum(x, -p.hi);
    r.lo = __dsub_rn(r.lo, p.lo);
    r = two_sum(r.hi, r.lo);
    dd q = two_prod(k, PIO2_2);
    r = dd_add(r, dd{-q.hi, -q.lo});
    r.lo = __dsub_rn(r.lo, __dmul_rn(k, PIO2_3));
    r = two_sum(r.hi, r.lo);
    const dd z = dd_mul(r, r);
    dd ps = {c_S[12][0], c_S[12][1]}, pc = {c_C[12][0], c_C[12][1]};
#pragma unroll 1
    for (int i = 11; i >= 0; --i) {
        ps = dd_add(dd_mul(ps, z), dd{c_S[i][0], c_S[i][1]});
        pc = dd_add(dd_mul(pc, z), dd{c_C[i][0], c_C[i][1]});
    }
    const dd sn = dd_add(r, dd_mul(dd_mul(r, z), ps));        // r + r^3 * S(z)
    const dd cs = dd_add(dd{1.0, 0.0}, dd_mul(z, pc));        // 1 + z * C(z)
    const long long kk = (long long)k;
    switch ((int)(kk & 3)) {
    case 0: *s_out = sn.hi; *c_out = cs.hi; break;
    case 1: *s_out = cs.hi; *c_out = -sn.hi; break;
    case 2: *s_out = -sn.hi; *c_out = -cs.hi; break;
    default: *s_out = -cs.hi; *c_out = sn.hi; break;
    }
}

// sin(x) (want_cos = 0) or cos(x) (want_cos = 1) alone: the same reduction and the same operations as sincos_cr for the
// one series the quadrant asks for (bit-identical to the corresponding output of sincos_cr), half the double-double work.
// The series is chosen per lane by selecting the coefficients, not by branching.
__device__ double sin_or_cos_cr(double x, int want_cos) {
    const double PIO2_1 = 1.5707963267948966, PIO2_2 = 6.123233995736766e-17, PIO2_3 = -1.4973849048591698e-33;
    const double k = rint(__dmul_rn(x, 0.6366197723675814));
    dd p = two_prod(k, PIO2_1);
    dd r = two_sum(x, -p.hi);
    r.lo = __dsub_rn(r.lo, p.lo);
    r = two_sum(r.hi, r.lo);
    dd q = two_prod(k, PIO2_2);
    r = dd_add(r, dd{-q.hi, -q.lo});
    r.lo = __dsub_rn(r.lo, __dmul_rn(k, PIO2_3));
    r = two_sum(r.hi, r.lo);
    const dd z = dd_mul(r, r);
    const int quad = (int)((long long)k & 3);
    const bool cos_series = ((quad ^ want_cos) & 1) != 0;
    // terms beyond z^(top+2) are below 2^-106 of the result for the given bound on z = r^2 (the DDS tables call this
    // almost only next to multiples of pi/2, where r ~ 1e-13 and two terms are the whole series)
    const double za = fabs(z.hi);
    const int top = za < 1e-16 ? 2 : (za < 1e-8 ? 4 : (za < 1e-3 ? 7 : (za < 0.05 ? 10 : 12)));
    dd ps = {cos_series ? c_C[top][0] : c_S[top][0], cos_series ? c_C[top][1] : c_S[top][1]};
#pragma unroll 1
    for (int i = top - 1; i >= 0; --i) {
        const double c0 = cos_series ? c_C[i][0] : c_S[i][0], c1 = cos_series ? c_C[i][1] : c_S[i][1];
        ps = dd_add(dd_mul(ps, z), dd{c0, c1});
    }
    const dd a = cos_series ? dd{1.0, 0.0} : r;                 // 1 + z * C(z)   |   r + (r * z) * S(z)
    const dd m = cos_series ? z : dd_mul(r, z);
    const double v = dd_add(a, dd_mul(m, ps)).hi;
    const bool neg = ((want_cos ? quad + 1 : quad) & 2) != 0;
    return neg ? -v : v;
}

// argument of the reference: ((2*pi*f)*(t+offset))/fs + phi, evaluated left to right in float64
// (ROACH_Setup.py:439-440)
__device__ __forceinline__ double ref_arg(double f, double t, double fs, double phi) {
    const double w = __dmul_rn(6.283185307179586, f);
    return __dadd_rn(__ddiv_rn(__dmul_rn(w, t), fs), phi);
}

__device__ __forceinline__ double ref_arg_w(double w, double t, double fs, double phi) {       // w = RN(2 pi f)
    return __dadd_rn(__ddiv_rn(__dmul_rn(w, t), fs), phi);
}

struct CombParams {
    const double *freq, *amp, *phase;    // [batch][T]
    long long *kbin;                     // [batch][T] spectral line index in [0, N) (written by comb_prep_kernel)
    unsigned int *bad;                   // set when a tone is not on the fs/N grid
    int T, N, N1, N2, offset;
    double fs;
    double2 *x;                          // [batch][N] bulk result
    unsigned long long *maxbits;         // [batch] bits of the bulk max(|I|,|Q|)
    double *scale;                       // [batch]
    unsigned long long *exact_max;       // [batch] bits of the exact max
    unsigned int *list;                  // [batch][cap] flagged samples: t | (isQ << 31)
    unsigned int *count;                 // [batch]
    unsigned int cap;
    double fudge, scale_override;
    int16_t *I, *Q;                      // [batch][N]
    const double2 *tw;                   // [N1] e^{+2 pi j k / N1} = (cos, sin): the twiddles of every radix-4 stage
    const double2 *tone;                 // [batch][T] (cos phi, sin phi) of every tone
    double *sigma;                       // [batch] estimate of the rms difference reference - bulk of one sample (signal units)
    double *eps;                         // [batch] flag distance of the quantiser (LSB)
    double *row_max;                     // [batch][N2] max(|I|,|Q|) of the N1 bulk samples of one n2 (one IFFT CTA)
};

// tables shared by all CTAs of comb_ifft_kernel (a double-precision sincospi costs ~100 instructions: computed per
// butterfly it was 60 % of that kernel's instruction stream).  tw holds, for every radix-4 stage Ns = 4, 16, ... < N1,
// the Ns factors e^{+2 pi j k / (4 Ns)}, k < Ns, at offset (Ns - 4) / 3: the lanes of a warp read CONSECUTIVE entries
// (a strided gather from one length-N1 table cost as many L1 wavefronts as all the shared-memory traffic of the FFT).
__global__ void comb_prep_kernel(CombParams p, int batch, double2 *tw, double2 *tone) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < (p.N1 - 4) / 3) {
        int Ns = 4;
        while ((4 * Ns - 4) / 3 <= i) Ns *= 4;
        const int k = i - (Ns - 4) / 3;
        double s, c;
        sincospi(2.0 * (double)k / (double)(4 * Ns), &s, &c);
        tw[i] = make_double2(c, s);
    }
    if (i < batch * p.T) {
        double sp, cp;
        sincos(p.phase[i], &sp, &cp);
        tone[i] = make_double2(cp, sp);
        // spectral line of the tone; must be on the fs/N grid (define_DAC_LUT snaps to it, :498)
        const double k = __ddiv_rn(__dmul_rn(p.freq[i], (double)p.N), p.fs);
        const double kr = rint(k);
        if (!(fabs(k - kr) < 1e-6)) atomicOr(p.bad, 1u);
        long long kk = (long long)kr % p.N;
        if (kk < 0) kk += p.N;
        p.kbin[i] = kk;
    }
    if (i < batch * 32) {
        // The reference rounds 2*pi*f, the product with t, the division by fs and the sum with phi: four roundings of an
        // argument of up to 2 pi N rad, each below 2^-53 relative, independent from tone to tone.  The bulk IFFT has none
        // of them (its own error is ~1e-15 of the scale), so reference - bulk of one sample is a sum of T terms
        // a_n * (argument error): rms below 2.2e-16 * 2 pi N * sqrt(sum a_n^2) (twice the rms of four uniform roundings
        // at the LARGEST argument).  Candidates for the max and samples near a truncation boundary are taken within
        // 8 of these sigmas and re-evaluated in the reference's order.  (One warp per LUT set: a single thread walking
        // the amplitudes made this 3 us kernel take 23 us.)
        const int bset = i >> 5, lane = i & 31;
        double ss = 0.0;
        for (int n = lane; n < p.T; n += 32) { const double a = p.amp[(size_t)bset * p.T + n]; ss += a * a; }
        for (int d = 16; d > 0; d >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, d);
        if (lane == 0) p.sigma[bset] = 2.2e-16 * 6.283185307179586 * (double)p.N * sqrt(ss);
    }
}

// shared-memory position of FFT element i: 16-byte elements are served per quarter-warp (8 lanes over the 8 four-bank
// groups), so the bank group is i & 7.  Stage Ns = 1 stores element 4 j + r and stage Ns = 4 element 16 q + k + 4 r from
// lane j = 4 q + k: 4-way and 2-way conflicts in the plain layout.  XOR-ing the group with bits of i >> 3 keeps every
// aligned run of 8 elements conflict-free (all loads, stores of the stages Ns >= 16) and spreads those two stores.
__device__ __forceinline__ int fft_swz(int i) {
    const int h = i >> 3;
    return i ^ ((h ^ (h << 1)) & 7);
}

__device__ __forceinline__ double2 zmul(double2 a, double2 w) { return make_double2(a.x * w.x - a.y * w.y, a.x * w.y + a.y * w.x); }
__device__ __forceinline__ void idft4(double2 &a0, double2 &a1, double2 &a2, double2 &a3) {     // X[m] = sum a[r] (+j)^(r m)
    const double2 t0 = make_double2(a0.x + a2.x, a0.y + a2.y), t1 = make_double2(a0.x - a2.x, a0.y - a2.y);
    const double2 t2 = make_double2(a1.x + a3.x, a1.y + a3.y), t3 = make_double2(-(a1.y - a3.y), a1.x - a3.x);
    a0 = make_double2(t0.x + t2.x, t0.y + t2.y); a1 = make_double2(t1.x + t3.x, t1.y + t3.y);
    a2 = make_double2(t0.x - t2.x, t0.y - t2.y); a3 = make_double2(t1.x - t3.x, t1.y - t3.y);
}

// ---- K1a: per (n2, batch) a length-N1 Stockham radix-4 IFFT in shared memory
template <int N1>
__global__ void __launch_bounds__(N1 / 4 < 32 ? 32 : N1 / 4) comb_ifft_kernel(CombParams p) {
    __shared__ double2 bufA[N1], bufB[N1];
    __shared__ double s_red[32];
    const int n2 = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
    constexpr int NT = N1 / 4;                       // butterflies per stage
    constexpr int BT = NT < 32 ? 32 : NT;            // threads per CTA: whole warps (the short tables leave lanes idle)
    for (int i = tid; i < N1; i += BT) bufA[i] = make_double2(0.0, 0.0);
    __syncthreads();
    // sparse fill: G[k mod N1] += a e^{j phi} e^{2 pi j k n2 / N}
    for (int i = tid; i < p.T; i += BT) {
        const long long k = p.kbin[(size_t)b * p.T + i];
        const long long m = (k * (long long)n2) & (long long)(p.N - 1);      // N is a power of two
        double s, c;
        sincospi(2.0 * (double)m / (double)p.N, &s, &c);
        const double2 tn = p.tone[(size_t)b * p.T + i];
        const double a = p.amp[(size_t)b * p.T + i];
        const double2 g = zmul(make_double2(a * tn.x, a * tn.y), make_double2(c, s));
        const int k1 = fft_swz((int)(k & (N1 - 1)));
        atomicAdd(&bufA[k1].x, g.x);
        atomicAdd(&bufA[k1].y, g.y);
    }
    __syncthreads();
    double2 *in = bufA, *out = bufB;
    double mx = 0.0;
#pragma unroll 1
    for (int Ns = 1; Ns < N1; Ns *= 4) {
        const bool last = Ns * 4 == N1;
        if (BT == NT || tid < NT) {
            const int j = tid, k = j % Ns;
            double2 v[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) v[r] = in[fft_swz(j + r * NT)];
            if (Ns > 1) {
                // the bulk result only has to be good to ~1e-9 relative (samples within eps LSB of a rounding boundary
                // are re-evaluated exactly): w^2 and w^3 by multiplication instead of two more table reads
                const double2 w1 = __ldg(&p.tw[(Ns - 4) / 3 + k]);              // e^{+2 pi j k/(4 Ns)}
                const double2 w2 = zmul(w1, w1), w3 = zmul(w2, w1);
                v[1] = zmul(v[1], w1); v[2] = zmul(v[2], w2); v[3] = zmul(v[3], w3);
            }
            idft4(v[0], v[1], v[2], v[3]);                                      // inverse radix-4 butterfly (twiddle +j)
            if (!last) {
                const int j0 = (j - k) * 4 + k;
#pragma unroll
                for (int r = 0; r < 4; ++r) out[fft_swz(j0 + r * Ns)] = v[r];
            } else {
                // last stage (k = j, j0 = j): straight from the registers to HBM.  Sample t = n2 + N2 * n1 is stored at
                // x[n2 * N1 + n1]: whole lines per warp (the time-ordered layout made every 16-byte store its own L1
                // wavefront); the consumers transpose (comb_quantise_kernel) or do not care about order
                double2 *x = p.x + (size_t)b * p.N + (size_t)n2 * N1 + j;
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    x[r * NT] = v[r];
                    mx = fmax(mx, fmax(fabs(v[r].x), fabs(v[r].y)));
                }
            }
        }
        if (!last) __syncthreads();
        double2 *tmp = in; in = out; out = tmp;
    }
    for (int d = 16; d > 0; d >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, d));
    if ((tid & 31) == 0) s_red[tid >> 5] = mx;
    __syncthreads();
    if (tid == 0) {
        for (int i = 1; i < BT / 32; ++i) mx = fmax(mx, s_red[i]);
        p.row_max[(size_t)b * p.N2 + n2] = mx;
        atomicMax(&p.maxbits[b], (unsigned long long)__double_as_longlong(mx));
    }
}

// ---- K1a for N1 = 1024: radices 16 x 16 x 4.  Two of the five radix-4 stages of comb_ifft_kernel are done in registers
// (a 16-point inverse DFT per thread), so an FFT makes three trips through shared memory instead of five: that kernel is
// bound by shared-memory wavefronts.  One CTA of 256 threads transforms four rows n2 (64 threads each) in place.
// in place: on return y[c + 4 d] = sum_r v[r] e^{+2 pi j r (c + 4 d) / 16} sits in v[4 c + d]
__device__ __forceinline__ void idft16(double2 (&v)[16]) {
#pragma unroll
    for (int bb = 0; bb < 4; ++bb) idft4(v[bb], v[4 + bb], v[8 + bb], v[12 + bb]);      // over a (r = 4 a + b): v[4 c + b] = U_b[c]
    const double c1 = 0.92387953251128674, s1 = 0.38268343236508977, r2 = 0.70710678118654752;
    // U_b[c] *= W^(b c), W = e^{+2 pi j / 16}
    v[5] = zmul(v[5], make_double2(c1, s1));                                              // c = 1, b = 1: W^1
    v[6] = make_double2((v[6].x - v[6].y) * r2, (v[6].x + v[6].y) * r2);                  // c = 1, b = 2: W^2
    v[7] = zmul(v[7], make_double2(s1, c1));                                              // c = 1, b = 3: W^3
    v[9] = make_double2((v[9].x - v[9].y) * r2, (v[9].x + v[9].y) * r2);                  // c = 2, b = 1: W^2
    v[10] = make_double2(-v[10].y, v[10].x);                                              // c = 2, b = 2: W^4 = +j
    v[11] = make_double2(-(v[11].x + v[11].y) * r2, (v[11].x - v[11].y) * r2);            // c = 2, b = 3: W^6
    v[13] = zmul(v[13], make_double2(s1, c1));                                            // c = 3, b = 1: W^3
    v[14] = make_double2(-(v[14].x + v[14].y) * r2, (v[14].x - v[14].y) * r2);            // c = 3, b = 2: W^6
    v[15] = zmul(v[15], make_double2(-c1, -s1));                                          // c = 3, b = 3: W^9
#pragma unroll
    for (int c = 0; c < 4; ++c) idft4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);   // over b: v[4 c + d] = y[c + 4 d]
}
// element i of an FFT in shared memory: bank group (i & 7) XOR bits 4..6 of i.  Aligned runs of 8 elements (all loads,
// the stores of the second stage) stay conflict-free and the first stage's stores at stride 16 spread over the groups.
__device__ __forceinline__ int swz16(int i) { return i ^ ((i >> 4) & 7); }

template <int R16_F>                                         // rows (FFTs) per CTA, 64 threads each
__global__ void __launch_bounds__(64 * R16_F, 768 / (64 * R16_F)) comb_ifft1024_r16_kernel(CombParams p) {
    extern __shared__ __align__(16) unsigned char smem_r16[];
    double2 *buf_all = reinterpret_cast<double2 *>(smem_r16);           // [R16_F][1024]
    __shared__ double s_red[8];
    constexpr int N1 = 1024;
    const int tid = threadIdx.x, b = blockIdx.y, f = tid >> 6, j = tid & 63;
    const int n2_0 = blockIdx.x * R16_F, n2 = n2_0 + f;
    double2 *buf = buf_all + f * N1;
#pragma unroll
    for (int i = 0; i < 16; ++i) buf_all[tid + 64 * R16_F * i] = make_double2(0.0, 0.0);
    __syncthreads();
    // sparse fill of the four rows: G[k mod N1] += a e^{j phi} e^{2 pi j k n2 / N}; consecutive rows differ by e^{2 pi j k / N}
    for (int i = tid; i < p.T; i += 64 * R16_F) {
        const long long k = p.kbin[(size_t)b * p.T + i];
        const long long m = (k * (long long)n2_0) & (long long)(p.N - 1);      // N is a power of two
        double s, c, ss = 0.0, cs = 1.0;
        sincospi(2.0 * (double)m / (double)p.N, &s, &c);
        if (R16_F > 1) sincospi(2.0 * (double)k / (double)p.N, &ss, &cs);
        const double2 tn = p.tone[(size_t)b * p.T + i];
        const double a = p.amp[(size_t)b * p.T + i];
        double2 g = zmul(make_double2(a * tn.x, a * tn.y), make_double2(c, s));
        const int k1 = swz16((int)(k & (N1 - 1)));
#pragma unroll
        for (int ff = 0; ff < R16_F; ++ff) {
            if (n2_0 + ff < p.N2) {
                atomicAdd(&buf_all[ff * N1 + k1].x, g.x);
                atomicAdd(&buf_all[ff * N1 + k1].y, g.y);
            }
            if (R16_F > 1) g = zmul(g, make_double2(cs, ss));
        }
    }
    __syncthreads();
    double2 v[16];
    // stage A: radix 16, Ns = 1 (no twiddles): element 16 j + m
#pragma unroll
    for (int r = 0; r < 16; ++r) v[r] = buf[swz16(j + 64 * r)];
    idft16(v);
    __syncthreads();
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int d = 0; d < 4; ++d) buf[swz16(16 * j + c + 4 * d)] = v[4 * c + d];
    __syncthreads();
    // stage B: radix 16, Ns = 16: v[r] *= e^{2 pi j r k / 256}, k = j mod 16; element (j - k) 16 + k + 16 m
    {
        const int k = j & 15;
#pragma unroll
        for (int r = 0; r < 16; ++r) v[r] = buf[swz16(j + 64 * r)];
        const double2 w1 = __ldg(&p.tw[(64 - 4) / 3 + k]);                   // table of stage Ns = 64: e^{2 pi j k / 256}
        double2 w = w1;
        v[1] = zmul(v[1], w);
#pragma unroll
        for (int r = 2; r < 16; ++r) { w = zmul(w, w1); v[r] = zmul(v[r], w); }
        idft16(v);
        __syncthreads();
        const int base = (j - k) * 16 + k;
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
            for (int d = 0; d < 4; ++d) buf[swz16(base + 16 * (c + 4 * d))] = v[4 * c + d];
    }
    __syncthreads();
    // stage C: radix 4, Ns = 256, four butterflies per thread, straight to HBM: x[n2 * N1 + jj + 256 m]
    double mx = 0.0;
    if (n2 < p.N2) {
        double2 *x = p.x + (size_t)b * p.N + (size_t)n2 * N1;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int jj = j + 64 * u;
            double2 a0 = buf[swz16(jj)], a1 = buf[swz16(jj + 256)], a2 = buf[swz16(jj + 512)], a3 = buf[swz16(jj + 768)];
            const double2 w1 = __ldg(&p.tw[(256 - 4) / 3 + jj]);            // e^{2 pi j jj / 1024}
            const double2 w2 = zmul(w1, w1), w3 = zmul(w2, w1);
            a1 = zmul(a1, w1); a2 = zmul(a2, w2); a3 = zmul(a3, w3);
            idft4(a0, a1, a2, a3);
            x[jj] = a0; x[jj + 256] = a1; x[jj + 512] = a2; x[jj + 768] = a3;
            mx = fmax(mx, fmax(fmax(fmax(fabs(a0.x), fabs(a0.y)), fmax(fabs(a1.x), fabs(a1.y))),
                               fmax(fmax(fabs(a2.x), fabs(a2.y)), fmax(fabs(a3.x), fabs(a3.y)))));
        }
    }
    for (int d = 16; d > 0; d >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, d));
    if ((tid & 31) == 0) s_red[tid >> 5] = mx;
    __syncthreads();
    if (j == 0 && n2 < p.N2) {
        mx = fmax(s_red[2 * f], s_red[2 * f + 1]);
        p.row_max[(size_t)b * p.N2 + n2] = mx;
        atomicMax(&p.maxbits[b], (unsigned long long)__double_as_longlong(mx));
    }
}

// ---- K1b: samples that can hold the max -> list.  One CTA per (n2, set): only the rows whose own maximum (left by
// the IFFT CTA) reaches the limit are read at all - one to three of the N2 rows of a set.  Entry n1 of row n2 is Q at
// t = n2 + N2 * n1 and I at t - offset (I[t] is the real part at t + offset, ROACH_Setup_DAC.py:419-420)
__global__ void __launch_bounds__(256) comb_max_candidates_kernel(CombParams p) {
    const int n2 = blockIdx.x, b = blockIdx.y;
    const double mx = __longlong_as_double((long long)p.maxbits[b]);
    const double lim = mx - fmax(1e-8 * mx, 8.0 * p.sigma[b]);
    if (p.row_max[(size_t)b * p.N2 + n2] < lim) return;
    const double2 *x = p.x + (size_t)b * p.N + (size_t)n2 * p.N1;
    for (int n1 = threadIdx.x; n1 < p.N1; n1 += blockDim.x) {
        const double2 v = x[n1];
        const bool hq = fabs(v.y) >= lim, hi = fabs(v.x) >= lim;
        if (hq | hi) {
            const unsigned t = (unsigned)n2 + (unsigned)p.N2 * (unsigned)n1;
            if (hq) {
                const unsigned pos = atomicAdd(&p.count[b], 1u);
                if (pos < p.cap) p.list[(size_t)b * p.cap + pos] = t | (1u << 31);
            }
            if (hi) {
                const unsigned pos = atomicAdd(&p.count[b], 1u);
                if (pos < p.cap) p.list[(size_t)b * p.cap + pos] = (t - (unsigned)p.offset) & (unsigned)(p.N - 1);
            }
        }
    }
}

// reference-order evaluation of one sample by the whole CTA: terms in parallel (T / 128 double-double sin or cos per
// thread: the latency of a call with few sets is the latency of ONE sample), sum sequentially in tone order
__device__ double exact_sample(const CombParams &p, int b, int t, int isQ, double *s_terms) {
    const double tt = isQ ? (double)t : (double)(t + p.offset);
    for (int i = threadIdx.x; i < p.T; i += blockDim.x) {
        const double arg = ref_arg(p.freq[(size_t)b * p.T + i], tt, p.fs, p.phase[(size_t)b * p.T + i]);
        s_terms[i] = __dmul_rn(p.amp[(size_t)b * p.T + i], sin_or_cos_cr(arg, !isQ));
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double acc = 0.0;
        for (int i = 0; i < p.T; ++i) acc = __dadd_rn(acc, s_terms[i]);
        s_terms[p.T] = acc;
    }
    __syncthreads();
    const double acc = s_terms[p.T];
    __syncthreads();                     // the next sample overwrites s_terms
    return acc;
}

// ---- K1c: exact max over the candidates (one CTA per candidate), then the scale
__global__ void __launch_bounds__(128) comb_exact_max_kernel(CombParams p) {
    extern __shared__ double s_dyn[];
    const int b = blockIdx.y;
    const unsigned n = min(p.count[b], p.cap);
    for (unsigned c = blockIdx.x; c < n; c += gridDim.x) {
        const unsigned e = p.list[(size_t)b * p.cap + c];
        const double v = fabs(exact_sample(p, b, (int)(e & 0x7fffffffu), (int)(e >> 31), s_dyn));
        if (threadIdx.x == 0) atomicMax(&p.exact_max[b], (unsigned long long)__double_as_longlong(v));
    }
}

__global__ void comb_scale_kernel(CombParams p, int batch) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= batch) return;
    const double a = __longlong_as_double((long long)p.exact_max[b]);
    double sc = a;                                             // scale_factor = a.max()        (:448-449)
    if (p.fudge != 1.0) sc = __dmul_rn(p.fudge, sc);           // scaleFudgeFactor*scale_factor (:453-455)
    if (p.scale_override > 0.0) sc = p.scale_override;         // keep-old / custom scale       (:456-459)
    p.scale[b] = sc;
    // flag distance in LSB: 8 sigma, at least 1e-7 (covers the bulk's own error)
    p.eps[b] = fmax(1e-7, 8.0 * p.sigma[b] * 32767.0 / sc);
    p.count[b] = 0;                                            // the list is reused by the quantiser
}

// ---- K1d: quantise; flag samples within eps of an integer for exact re-evaluation.  One CTA per tile of QT_A values of
// n2 x QT_C values of n1: reads runs of QT_C entries of x (storage order), transposes the int16 results through shared
// memory and stores runs of QT_A consecutive samples.  Unflagged samples are further than eps from a rounding boundary,
// so x * (32767 / scale) truncates to the same integer as the reference's (x * 32767) / scale; the flagged ones are
// overwritten by comb_fixup_kernel with the reference-order value.
constexpr int QT_A = 64, QT_C = 32;
__global__ void __launch_bounds__(256) comb_quantise_kernel(CombParams p) {
    __shared__ int16_t sI[QT_C][QT_A + 2], sQ[QT_C][QT_A + 2];
    const int b = blockIdx.y, tid = threadIdx.x;
    const int tiles_c = (p.N1 + QT_C - 1) / QT_C;
    const int a0 = (blockIdx.x / tiles_c) * QT_A, c0 = (blockIdx.x % tiles_c) * QT_C;
    const double rs = __ddiv_rn(32767.0, p.scale[b]);
    const double eps = p.eps[b];
    const double2 *x = p.x + (size_t)b * p.N;
    const unsigned nmask = (unsigned)(p.N - 1);
#pragma unroll
    for (int i = 0; i < QT_A * QT_C / 256; ++i) {
        const int e = tid + i * 256, a = e / QT_C, c = e % QT_C;
        if (a0 + a < p.N2 && c0 + c < p.N1) {
            const double2 v = x[(size_t)(a0 + a) * p.N1 + c0 + c];
            const double vi = v.x * rs, vq = v.y * rs;
            const bool fi = fabs(vi - rint(vi)) < eps, fq = fabs(vq - rint(vq)) < eps;
            if (fi | fq) {
                const unsigned t = (unsigned)(a0 + a) + (unsigned)p.N2 * (unsigned)(c0 + c);
                if (fq) {
                    const unsigned pos = atomicAdd(&p.count[b], 1u);
                    if (pos < p.cap) p.list[(size_t)b * p.cap + pos] = t | (1u << 31);
                }
                if (fi) {
                    const unsigned pos = atomicAdd(&p.count[b], 1u);
                    if (pos < p.cap) p.list[(size_t)b * p.cap + pos] = (t - (unsigned)p.offset) & nmask;
                }
            }
            sI[c][a] = (int16_t)max(-32768, min(32767, __double2int_rz(vi)));
            sQ[c][a] = (int16_t)max(-32768, min(32767, __double2int_rz(vq)));
        }
    }
    __syncthreads();
    int16_t *Io = p.I + (size_t)b * p.N, *Qo = p.Q + (size_t)b * p.N;
#pragma unroll
    for (int i = 0; i < QT_A * QT_C / 256; ++i) {
        const int e = tid + i * 256, c = e / QT_A, a = e % QT_A;
        if (a0 + a < p.N2 && c0 + c < p.N1) {
            const unsigned t = (unsigned)(a0 + a) + (unsigned)p.N2 * (unsigned)(c0 + c);
            Qo[t] = sQ[c][a];
            Io[(t - (unsigned)p.offset) & nmask] = sI[c][a];
        }
    }
}

// ---- K1e: exact re-evaluation of the flagged samples
__global__ void __launch_bounds__(128) comb_fixup_kernel(CombParams p) {
    extern __shared__ double s_dyn[];
    const int b = blockIdx.y;
    const unsigned n = min(p.count[b], p.cap);
    const double sc = p.scale[b];
    for (unsigned c = blockIdx.x; c < n; c += gridDim.x) {
        const unsigned e = p.list[(size_t)b * p.cap + c];
        const int t = (int)(e & 0x7fffffffu), isQ = (int)(e >> 31);
        const double x = exact_sample(p, b, t, isQ, s_dyn);
        const double v = __ddiv_rn(__dmul_rn(x, 32767.0), sc);           // int(i*amp_full_scale/scale_factor) (:461-462)
        if (threadIdx.x == 0)
            (isQ ? p.Q : p.I)[(size_t)b * p.N + t] = (int16_t)max(-32768, min(32767, __double2int_rz(v)));
    }
}

// ---- K2: DDS tables, one CTA per (channel, batch).  The argument of every sample is the reference's own float64
// expression (ref_arg), so only sin / cos decide the result.  Bulk: the CUDA double-precision sincos (<= 2 ulp, i.e.
// <= 1.5e-11 LSB after scaling) for all samples; the correctly rounded double-double series (13 times the work) only
// where it can matter: the samples that can hold the table's maximum (its exact value is the scale) and the samples
// whose scaled value is within 1e-9 LSB of a truncation boundary.  The residuals are multiples of fs2/size, so sin and
// cos hit 0 and +-1 often (~1 % of the samples are of either kind, thousands in a channel whose residual is a multiple
// of a high power of two): they are first collected in a shared-memory list and then evaluated by all threads of the
// CTA, not in divergent branches.
// Fast bulk evaluation (round 2).  For a residual on the fs2/size grid, f = k fs2/size, the argument is
//     A = 2 pi (k t)/size + ph + dl,     dl = the rounding of the reference's own expression (a few 1e-12 rad),
// so sin A = sin(theta_j + ph + dl) with j = k t mod size: theta_j comes from a table of size entries (sincospi, shared
// by all channels), ph from ONE sincos per channel, and dl -- computed in extended precision from the reference's A
// itself: dl = ((A - hi) - ph) - lo with hi + lo = (k t)(2 pi / size) as a double-double -- enters to first order
// (dl^2 / 2 < 1e-16 is enforced; a sample that violates it, e.g. an off-grid residual, takes the library sincos).  Error
// of the bulk value <= 6e-16 = 2e-11 LSB after scaling, far inside the 1e-9 LSB flag distance of the exact path, which is
// unchanged; about 25 instead of about 110 fp64 operations per sample.
__device__ __forceinline__ int dds_swz(int j) { return j ^ ((j >> 3) & 7) ^ ((j >> 6) & 7) ^ ((j >> 9) & 7); }
__global__ void dds_table_kernel(double2 *tab, int size) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= size) return;
    double sn, cs;
    sincospi(2.0 * (double)j / (double)size, &sn, &cs);          // size is a power of two: the argument is exact
    tab[dds_swz(j)] = make_double2(sn, cs);
}
struct DdsFast {
    const double2 *tab; int mask; double c_hi, c_lo, sph, cph, ph, w, fs2, rfs; int k; bool on;
    __device__ __forceinline__ void init(const double2 *table, int size, double f, double phi, double fs) {
        tab = table; mask = size - 1; ph = phi; fs2 = fs; rfs = __ddiv_rn(1.0, fs);
        c_hi = 6.283185307179586 / (double)size; c_lo = 2.4492935982947064e-16 / (double)size;   // 2 pi = hi + lo
        w = __dmul_rn(6.283185307179586, f);
        const double kf = rint(f * (double)size / fs);
        on = table != nullptr && fabs(phi) < 64.0 && fabs(kf) <= (double)size;      // |k t| < 2^31
        k = on ? (int)kf : 0;
        sincos(phi, &sph, &cph);
    }
    // sin and cos of the reference's argument ((2 pi f) t) / fs2 + ph for the integer sample index t
    __device__ __forceinline__ void eval(int t, double *sn, double *cs) const {
        if (on) {
            // A exactly as ref_arg: the correctly rounded quotient x / fs2 from the reciprocal (q0 = x r faithful, one
            // exact-residual correction: Markstein's theorem, r = RN(1 / fs2))
            const double x = __dmul_rn(w, (double)t);
            const double q0 = __dmul_rn(x, rfs);
            const double q = __fma_rn(__fma_rn(-q0, fs2, x), rfs, q0);
            const double A = __dadd_rn(q, ph);
            const int kt = k * t;
            const double y = (double)kt;
            const double p_hi = __dmul_rn(y, c_hi);
            const double p_lo = __fma_rn(y, c_hi, -p_hi) + y * c_lo;
            const double dl = ((A - p_hi) - ph) - p_lo;
            const double2 e = __ldg(&tab[dds_swz(kt & mask)]);
            const double S0 = fma(e.x, cph, e.y * sph), C0 = fma(e.y, cph, -(e.x * sph));
            if (fabs(dl) < 1e-8) { *sn = fma(dl, C0, S0); *cs = fma(-dl, S0, C0); return; }
        }
        sincos(ref_arg_w(w, (double)t, fs2, ph), sn, cs);
    }
};

__global__ void __launch_bounds__(256) dds_lut_kernel(const double *resid, const double *phase, double fs2, int size,
                                                      int offset, const double2 *g_tab, int cpc, int16_t *I_dds, int16_t *Q_dds,
                                                      double *scales) {
    extern __shared__ double s_dyn[];            // I[size] | Q[size] | list[2 * size] (uint16: t | isI << 15) | exact bits[2 * size]
    __shared__ double s_red[8];
    __shared__ unsigned long long s_max;
    __shared__ unsigned s_n;
    double *sI = s_dyn, *sQ = s_dyn + size;
    unsigned short *list = reinterpret_cast<unsigned short *>(s_dyn + 2 * size);
    // bit e of s_exact: the staged value of entry e (t | isI << 15 -> bit t + isI * size) is already the correctly rounded
    // one (a candidate for the maximum, evaluated in the first exact round): the quantiser uses it as it is
    unsigned *s_exact = reinterpret_cast<unsigned *>(list + 2 * size);
    const int b = blockIdx.y, tid = threadIdx.x;
    for (int ci = 0; ci < cpc; ++ci) {
    const int m = blockIdx.x * cpc + ci;
    const double f = resid[(size_t)b * 256 + m], ph = phase[(size_t)b * 256 + m];
    DdsFast fast;
    fast.init(g_tab, size, f, ph, fs2);
    if (tid == 0) { s_max = 0ull; s_n = 0u; }
    for (int i = tid; i < (2 * size + 31) / 32; i += 256) s_exact[i] = 0u;
    __syncthreads();                              // the previous channel's lists are done
    double mx = 0.0;
    for (int t = tid; t < size; t += 256) {
        double s, c;
        fast.eval(t, &s, &c);
        sQ[t] = s;                                // amplitude 1., accumulated onto 0. : exact
        if (offset != 0) { double s2; fast.eval(t + offset, &s2, &c); }
        sI[t] = c;
        mx = fmax(mx, fmax(fabs(c), fabs(s)));
    }
    for (int d = 16; d > 0; d >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, d));
    if ((tid & 31) == 0) s_red[tid >> 5] = mx;
    __syncthreads();
    mx = s_red[0];
    for (int i = 1; i < 8; ++i) mx = fmax(mx, s_red[i]);
    // exact maximum: every sample within 1e-13 of the bulk maximum, correctly rounded (bit patterns of non-negative
    // doubles order like the values)
    const double lim = mx * (1.0 - 1e-13);
    for (int t = tid; t < size; t += 256) {
        if (fabs(sQ[t]) >= lim) list[atomicAdd(&s_n, 1u)] = (unsigned short)t;
        if (fabs(sI[t]) >= lim) list[atomicAdd(&s_n, 1u)] = (unsigned short)(t | 0x8000);
    }
    __syncthreads();
    for (unsigned i = tid; i < s_n; i += 256) {
        const int e = list[i], t = e & 0x7fff, isI = e >> 15;
        const double v = sin_or_cos_cr(ref_arg(f, (double)(isI ? t + offset : t), fs2, ph), isI);
        (isI ? sI : sQ)[t] = v;
        atomicOr(&s_exact[(t + isI * size) >> 5], 1u << ((t + isI * size) & 31));
        atomicMax(&s_max, (unsigned long long)__double_as_longlong(fabs(v)));
    }
    __syncthreads();
    mx = __longlong_as_double((long long)s_max);
    __syncthreads();                              // everybody has read s_max / s_n
    if (tid == 0) { if (scales) scales[(size_t)b * 256 + m] = mx; s_n = 0u; }
    __syncthreads();
    // channel-major tables [b][m][t] (whole lines per warp); dds_interleave_kernel scatters them to the reference's layout
    int16_t *Io = I_dds + ((size_t)b * 256 + m) * size, *Qo = Q_dds + ((size_t)b * 256 + m) * size;
    // a sample that is not flagged is further than 1e-9 from an integer: x * (32767 / mx) truncates like the reference's
    // (x * 32767) / mx
    const double rs = __ddiv_rn(32767.0, mx);
    for (int t = tid; t < size; t += 256) {
        const int dst = t;
        const double vi = sI[t] * rs, vq = sQ[t] * rs;
        // (a flagged sample that rounds to 0 truncates to 0 from either side: no exact evaluation)
        const double ri = rint(vi), rq = rint(vq);
        const unsigned eq = (unsigned)t, ei = (unsigned)(t + size);         // bit indices of the Q and I sample
        if (fabs(vi - ri) < 1e-9 && ri != 0.0) {
            if ((s_exact[ei >> 5] >> (ei & 31)) & 1u) Io[dst] = (int16_t)__double2int_rz(__ddiv_rn(__dmul_rn(sI[t], 32767.0), mx));
            else list[atomicAdd(&s_n, 1u)] = (unsigned short)(t | 0x8000);
        } else Io[dst] = (int16_t)__double2int_rz(vi);
        if (fabs(vq - rq) < 1e-9 && rq != 0.0) {
            if ((s_exact[eq >> 5] >> (eq & 31)) & 1u) Qo[dst] = (int16_t)__double2int_rz(__ddiv_rn(__dmul_rn(sQ[t], 32767.0), mx));
            else list[atomicAdd(&s_n, 1u)] = (unsigned short)t;
        } else Qo[dst] = (int16_t)__double2int_rz(vq);
    }
    __syncthreads();
    for (unsigned i = tid; i < s_n; i += 256) {
        const int e = list[i], t = e & 0x7fff, isI = e >> 15;
        const int dst = t;
        const double x = sin_or_cos_cr(ref_arg(f, (double)(isI ? t + offset : t), fs2, ph), isI);
        (isI ? Io : Qo)[dst] = (int16_t)__double2int_rz(__ddiv_rn(__dmul_rn(x, 32767.0), mx));
    }
    __syncthreads();                              // the lists and s_n are reused by the next channel
    }
}

// ---- K2b: channel-major pairs [b][m][j] (samples 2j, 2j+1 of channel m as one 32-bit word) -> the reference's layout
// dds[j*512 + 2*((m+ch_shift)%256) + s] (ROACH_Setup.py:526-530): a 32-bit transpose with a rotation of the channel
// axis, tiles of 32 x 32 words through shared memory.  (Stored straight from dds_lut_kernel, every 2-byte sample was its
// own 32-byte sector: 1 KiB stride.)
__global__ void __launch_bounds__(256) dds_interleave_kernel(const uint32_t *I_cm, const uint32_t *Q_cm, int half, int ch_shift,
                                                             uint32_t *I_dds, uint32_t *Q_dds) {
    __shared__ uint32_t tI[32][33], tQ[32][33];
    const int b = blockIdx.y, tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int m0 = (blockIdx.x & 7) * 32, j0 = (blockIdx.x >> 3) * 32;
    const size_t base = (size_t)b * 256 * half;
    if (j0 + tx < half) {
#pragma unroll
        for (int r = ty; r < 32; r += 8) {
            tI[r][tx] = I_cm[base + (size_t)(m0 + r) * half + j0 + tx];
            tQ[r][tx] = Q_cm[base + (size_t)(m0 + r) * half + j0 + tx];
        }
    }
    __syncthreads();
    const int slot = (m0 + tx + ch_shift) & 255;
#pragma unroll
    for (int r = ty; r < 32; r += 8) {
        if (j0 + r < half) {
            I_dds[base + (size_t)(j0 + r) * 256 + slot] = tI[tx][r];
            Q_dds[base + (size_t)(j0 + r) * 256 + slot] = tQ[tx][r];
        }
    }
}

// ---- K3: DRAM image, 16 bytes per sample pair: >h of q_dds1 q_dds0 q_dac1 q_dac0 i_dds1 i_dds0 i_dac1 i_dac0
__device__ __forceinline__ uint32_t be_pair(int16_t first, int16_t second) {
    // bytes in memory: first.hi first.lo second.hi second.lo  (little-endian u32 store)
    const uint32_t a = (uint16_t)first, c = (uint16_t)second;
    return (a >> 8) | ((a & 0xFF) << 8) | ((c >> 8) << 16) | ((c & 0xFF) << 24);
}
__global__ void pack_dram_kernel(const int16_t *I_dac, const int16_t *Q_dac, const int16_t *I_dds, const int16_t *Q_dds,
                                 int64_t n_pairs, uint4 *out) {
    for (int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; n < n_pairs; n += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t qd = reinterpret_cast<const uint32_t *>(Q_dds)[n], qa = reinterpret_cast<const uint32_t *>(Q_dac)[n];
        const uint32_t id = reinterpret_cast<const uint32_t *>(I_dds)[n], ia = reinterpret_cast<const uint32_t *>(I_dac)[n];
        uint4 o;
        o.x = be_pair((int16_t)(qd >> 16), (int16_t)(qd & 0xFFFF));
        o.y = be_pair((int16_t)(qa >> 16), (int16_t)(qa & 0xFFFF));
        o.z = be_pair((int16_t)(id >> 16), (int16_t)(id & 0xFFFF));
        o.w = be_pair((int16_t)(ia >> 16), (int16_t)(ia & 0xFFFF));
        out[n] = o;
    }
}

__global__ void sincos_cr_test_kernel(const double *x, int64_t n, double *s, double *c) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (i & 1) { s[i] = sin_or_cos_cr(x[i], 0); c[i] = sin_or_cos_cr(x[i], 1); }    // both evaluators are under test
    else sincos_cr(x[i], &s[i], &c[i]);
}

// ------------------------------------------------------------------ MT19937 (numpy.random.seed / uniform)
struct MT19937 {
    uint32_t mt[624]; int idx;
    explicit MT19937(uint32_t seed) {            // init_genrand
        mt[0] = seed;
        for (int i = 1; i < 624; ++i) mt[i] = 1812433253u * (mt[i - 1] ^ (mt[i - 1] >> 30)) + (uint32_t)i;
        idx = 624;
    }
    uint32_t next() {
        if (idx >= 624) {
            for (int k = 0; k < 624; ++k) {
                uint32_t y = (mt[k] & 0x80000000u) | (mt[(k + 1) % 624] & 0x7fffffffu);
                mt[k] = mt[(k + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
            }
            idx = 0;
        }
        uint32_t y = mt[idx++];
        y ^= y >> 11; y ^= (y << 7) & 0x9d2c5680u; y ^= (y << 15) & 0xefc60000u; y ^= y >> 18;
        return y;
    }
    double res53() { uint32_t a = next() >> 5, b = next() >> 6; return (a * 67108864.0 + b) / 9007199254740992.0; }
};

}  // namespace

// numpy.random.seed(1000); phase[n] = numpy.random.uniform(0, 2*numpy.pi)  (ROACH_Setup.py:426-429)
extern "C" int mkid_random_phases(uint32_t seed, int32_t n, double *out) {
    if (!out || n < 0) return MKID_EINVAL;
    MT19937 g(seed);
    const double two_pi = 2 * 3.141592653589793;
    for (int i = 0; i < n; ++i) out[i] = 0.0 + (two_pi - 0.0) * g.res53();
    return MKID_OK;
}

extern "C" int mkid_comb_lut(mkid_ctx *ctx, const double *freq_hz, const double *amp, double *phase, int32_t n_tones,
                             double sample_rate, int32_t n_samples, int32_t offset, double fudge, int32_t random_phase,
                             double scale_override, int32_t batch, int16_t *I, int16_t *Q, double *scale_out) {
    if (!ctx) return MKID_EINVAL;
    MKID_REQUIRE(ctx, freq_hz && amp && phase && I && Q, "comb_lut: NULL argument");
    MKID_REQUIRE(ctx, n_tones >= 1 && n_tones <= 1024 && batch >= 1, "comb_lut: 1..1024 tones, batch >= 1");
    const int N = n_samples;
    MKID_REQUIRE(ctx, N >= 64 && N <= (1 << 24) && (N & (N - 1)) == 0, "comb_lut: n_samples must be a power of two in [64, 2^24]");
    int N1 = 1024;                         // largest power of four <= min(1024, N/2)
    while (N1 * 2 > N) N1 /= 4;
    const int N2 = N / N1;
    const size_t TB = (size_t)batch * n_tones;
    // every tone must sit on the fs/N grid (define_DAC_LUT snaps to it, :498): checked here, in the arithmetic of the device
    // check in comb_prep_kernel, BEFORE any work is queued, so that a bad call leaves the caller's I / Q untouched
    for (size_t i = 0; i < TB; ++i) {
        const double k = freq_hz[i] * (double)N / sample_rate;
        if (!(fabs(k - rint(k)) < 1e-6))
            return mkid_fail(ctx, MKID_EINVAL, "comb_lut: tone frequency is not a multiple of sampleRate/n_samples (set %d, tone %d: %.17g Hz)",
                             (int)(i / n_tones), (int)(i % n_tones), freq_hz[i]);
    }
    MKID_CUDA(ctx, cudaSetDevice(ctx->device));
    // one host block in the device layout freq | amp | phase: a single upload (the spectral lines are derived on the GPU)
    std::vector<double> hmeta(TB * 3);
    double *ph = hmeta.data() + 2 * TB;
    memcpy(hmeta.data(), freq_hz, TB * 8);
    memcpy(hmeta.data() + TB, amp, TB * 8);
    memcpy(ph, phase, TB * 8);
    if (random_phase) {                                       // every call re-seeds (:426): the same draws for every set
        mkid_random_phases(1000u, n_tones, ph);
        for (int b = 1; b < batch; ++b) memcpy(ph + (size_t)b * n_tones, ph, (size_t)n_tones * 8);
    }
    if (random_phase) memcpy(phase, ph, TB * 8);
    // device buffers
    // list of samples for the exact path, per set: ~100 (max candidates: a handful) at 2^19 samples and 256 tones; the flag
    // distance grows with N * sqrt(sum a^2), so the longest tables get a longer list
    const unsigned cap = std::max(1u << 17, (unsigned)N >> 5);
    char *meta; double2 *x; unsigned *list;
    int rc;
    const size_t meta_bytes = TB * 32 + (size_t)batch * 48;
    if ((rc = mkid_scratch(ctx, SCR_META, meta_bytes, (void **)&meta))) return rc;
    int sub = (int)std::max<size_t>(1, ((size_t)512 << 20) / ((size_t)N * 16));
    if (const char *e = getenv("MKID_LUT_GROUP")) sub = std::max(1, atoi(e));      // experiment switch
    bool radix16 = true;
    if (const char *e = getenv("MKID_LUT_RADIX4")) radix16 = atoi(e) == 0;           // experiment switch: five radix-4 stages
    int r16_f = 4;
    if (const char *e = getenv("MKID_LUT_R16F")) r16_f = atoi(e);                     // experiment switch: rows per CTA (1, 2, 4)
    if (radix16 && N1 == 1024) {
        static bool attr_set[64] = {};
        if (!attr_set[ctx->device & 63]) {
            MKID_CUDA(ctx, cudaFuncSetAttribute(comb_ifft1024_r16_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 1024 * 16));
            attr_set[ctx->device & 63] = true;
        }
    }
    if ((rc = mkid_scratch(ctx, SCR_AUX0, (size_t)sub * N * 16, (void **)&x))) return rc;
    if ((rc = mkid_scratch(ctx, SCR_AUX1, (size_t)batch * cap * 4, (void **)&list))) return rc;
    CombParams p;
    double *d_freq = (double *)meta, *d_amp = d_freq + TB, *d_phase = d_amp + TB;
    long long *d_k = (long long *)(d_phase + TB);
    unsigned long long *d_max = (unsigned long long *)(d_k + TB), *d_emax = d_max + batch;
    double *d_scale = (double *)(d_emax + batch);
    unsigned *d_count = (unsigned *)(d_scale + batch);
    MKID_CUDA(ctx, cudaMemcpyAsync(d_freq, hmeta.data(), TB * 24, cudaMemcpyHostToDevice, ctx->stream));
    MKID_CUDA(ctx, cudaMemsetAsync(d_max, 0, (size_t)batch * 48 - 0, ctx->stream));
    void *dI, *dQ;
    if ((rc = mkid_stage_out(ctx, I, (size_t)batch * N * 2, SCR_OUT0, false, &dI))) return rc;
    if ((rc = mkid_stage_out(ctx, Q, (size_t)batch * N * 2, SCR_OUT1, false, &dQ))) return rc;
    p.freq = d_freq; p.amp = d_amp; p.phase = d_phase; p.kbin = d_k; p.T = n_tones; p.N = N; p.N1 = N1; p.N2 = N2; p.offset = offset;
    p.fs = sample_rate; p.x = x; p.maxbits = d_max; p.scale = d_scale; p.exact_max = d_emax; p.list = list;
    p.count = d_count; p.bad = d_count + batch; p.cap = cap; p.fudge = fudge; p.scale_override = scale_override; p.I = (int16_t *)dI; p.Q = (int16_t *)dQ;
    double2 *d_tw;
    if ((rc = mkid_scratch(ctx, SCR_AUX2, ((size_t)N1 + TB) * 16 + (size_t)batch * (N2 + 2) * 8, (void **)&d_tw))) return rc;
    p.tw = d_tw; p.tone = d_tw + N1; p.sigma = (double *)(d_tw + N1 + TB); p.eps = p.sigma + batch; p.row_max = p.eps + batch;
    {
        const int n_prep = (int)std::max<size_t>(std::max<size_t>((size_t)N1, TB), (size_t)batch * 32);      // (N1 - 4) / 3 table entries, TB tones, a warp per set
        comb_prep_kernel<<<(n_prep + 255) / 256, 256, 0, ctx->stream>>>(p, batch, d_tw, d_tw + N1);
        MKID_CHECK_LAUNCH(ctx);
    }
    const int q_tiles = ((N2 + QT_A - 1) / QT_A) * ((N1 + QT_C - 1) / QT_C);
    const size_t wsm = (size_t)(n_tones + 1) * 8;
    const int fix_ctas = batch >= 16 ? 64 : 256;             // few sets: spread the ~100 flagged samples of a set
    // groups of `sub` LUT sets share ONE bulk buffer of <= 512 MiB (64 sets of 2^19 samples), so a large batch does not
    // scale the scratch.  Measured: L2-sized groups (4 / 8 / 16 sets) are SLOWER, 40 / 54.5 / 55.4 k sets per second
    // against 74.3 k for groups of 64: short launches pay their tails, the L2 hits do not make up for it.
    for (int b0 = 0; b0 < batch; b0 += sub) {
        const int nb = std::min(sub, batch - b0);
        CombParams q = p;
        const size_t t0 = (size_t)b0 * n_tones;
        q.freq += t0; q.amp += t0; q.phase += t0; q.kbin += t0; q.tone += t0;
        q.row_max += (size_t)b0 * N2; q.sigma += b0; q.eps += b0; q.maxbits += b0; q.scale += b0; q.exact_max += b0; q.count += b0; q.list += (size_t)b0 * cap;
        q.I += (size_t)b0 * N; q.Q += (size_t)b0 * N;
        dim3 g1(N2, nb);
        switch (N1) {
        case 16: comb_ifft_kernel<16><<<g1, 32, 0, ctx->stream>>>(q); break;
        case 64: comb_ifft_kernel<64><<<g1, 32, 0, ctx->stream>>>(q); break;
        case 256: comb_ifft_kernel<256><<<g1, 64, 0, ctx->stream>>>(q); break;
        default:
            if (radix16 && r16_f == 1) comb_ifft1024_r16_kernel<1><<<dim3(N2, nb), 64, 1024 * 16, ctx->stream>>>(q);
            else if (radix16 && r16_f == 2) comb_ifft1024_r16_kernel<2><<<dim3((N2 + 1) / 2, nb), 128, 2 * 1024 * 16, ctx->stream>>>(q);
            else if (radix16) comb_ifft1024_r16_kernel<4><<<dim3((N2 + 3) / 4, nb), 256, 4 * 1024 * 16, ctx->stream>>>(q);
            else comb_ifft_kernel<1024><<<g1, 256, 0, ctx->stream>>>(q);
            break;
        }
        MKID_CHECK_LAUNCH(ctx);
        if (scale_override <= 0.0) {
            comb_max_candidates_kernel<<<dim3(N2, nb), 256, 0, ctx->stream>>>(q);
            MKID_CHECK_LAUNCH(ctx);
            comb_exact_max_kernel<<<dim3(16, nb), 128, wsm, ctx->stream>>>(q);
            MKID_CHECK_LAUNCH(ctx);
        }
        comb_scale_kernel<<<(nb + 63) / 64, 64, 0, ctx->stream>>>(q, nb);
        MKID_CHECK_LAUNCH(ctx);
        comb_quantise_kernel<<<dim3(q_tiles, nb), 256, 0, ctx->stream>>>(q);
        MKID_CHECK_LAUNCH(ctx);
        comb_fixup_kernel<<<dim3(fix_ctas, nb), 128, wsm, ctx->stream>>>(q);
        MKID_CHECK_LAUNCH(ctx);
    }
    // overflow of the candidate / fix-up list would silently skip exact re-evaluation: check
    std::vector<double> back((size_t)batch + (batch + 2) / 2);          // scale[batch] | count[batch] | bad: adjacent on the device
    MKID_CUDA(ctx, cudaMemcpyAsync(back.data(), d_scale, (size_t)batch * 12 + 4, cudaMemcpyDeviceToHost, ctx->stream));
    MKID_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    const double *sc = back.data();
    const unsigned *cnt = reinterpret_cast<const unsigned *>(back.data() + batch);
    MKID_REQUIRE(ctx, cnt[batch] == 0, "comb_lut: tone frequency is not a multiple of sampleRate/n_samples");
    for (int b = 0; b < batch; ++b) {
        if (cnt[b] > cap) return mkid_fail(ctx, MKID_EINVAL, "comb_lut: %u samples need exact re-evaluation (> %u)", cnt[b], cap);
        if (scale_out) scale_out[b] = sc[b];
    }
    if ((rc = mkid_stage_out_finish(ctx, I, (size_t)batch * N * 2, dI))) return rc;
    return mkid_stage_out_finish(ctx, Q, (size_t)batch * N * 2, dQ);
}

extern "C" int mkid_dds_lut(mkid_ctx *ctx, const double *resid_hz, const double *phase, double sample_rate, int32_t n_lut,
                            int32_t ch_shift, int32_t offset, int32_t batch, int16_t *I_dds, int16_t *Q_dds,
                            double *scales_out) {
    if (!ctx) return MKID_EINVAL;
    MKID_REQUIRE(ctx, resid_hz && phase && I_dds && Q_dds && batch >= 1, "dds_lut: NULL argument");
    MKID_REQUIRE(ctx, n_lut >= 512 && n_lut % 512 == 0, "dds_lut: n_lut must be a multiple of 512");
    MKID_CUDA(ctx, cudaSetDevice(ctx->device));
    const double fs2 = sample_rate / 512.0 * 2.0;              // sampleRate/fft_len*2   (ROACH_Setup.py:525)
    const double res = sample_rate / (double)n_lut;
    const int size = (int)(fs2 / res);                          // size = int(sampleRate/resolution) (:422)
    MKID_REQUIRE(ctx, size * 256 == n_lut, "dds_lut: table size mismatch");
    int rc;
    double *meta;
    if ((rc = mkid_scratch(ctx, SCR_META, (size_t)batch * 256 * 24, (void **)&meta))) return rc;
    double *d_res = meta, *d_ph = meta + (size_t)batch * 256, *d_sc = d_ph + (size_t)batch * 256;
    MKID_CUDA(ctx, cudaMemcpyAsync(d_res, resid_hz, (size_t)batch * 256 * 8, cudaMemcpyDefault, ctx->stream));
    MKID_CUDA(ctx, cudaMemcpyAsync(d_ph, phase, (size_t)batch * 256 * 8, cudaMemcpyDefault, ctx->stream));
    void *dI, *dQ;
    if ((rc = mkid_stage_out(ctx, I_dds, (size_t)batch * n_lut * 2, SCR_OUT0, false, &dI))) return rc;
    if ((rc = mkid_stage_out(ctx, Q_dds, (size_t)batch * n_lut * 2, SCR_OUT1, false, &dQ))) return rc;
    size_t smem = (size_t)size * 20 + (size_t)((2 * size + 31) / 32) * 4;      // I, Q (fp64), the list of samples for the exact path, their bit map
    MKID_REQUIRE(ctx, smem <= 200 * 1024 && size <= 32768, "dds_lut: table too long for shared memory");
    // fast bulk path: sin / cos of the size grid angles from a shared-memory table (power-of-two sizes that fit)
    bool fast = size >= 64 && (size & (size - 1)) == 0;
    if (const char *e = getenv("MKID_DDS_SLOW")) fast = fast && atoi(e) == 0;      // experiment switch: library sincos for every sample
    double2 *d_tab = nullptr;
    if (fast) {
        if ((rc = mkid_scratch(ctx, SCR_AUX2, (size_t)size * 16, (void **)&d_tab))) return rc;
        dds_table_kernel<<<(size + 255) / 256, 256, 0, ctx->stream>>>(d_tab, size);
        MKID_CHECK_LAUNCH(ctx);
    }
    MKID_CUDA(ctx, cudaFuncSetAttribute(dds_lut_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int16_t *cmI, *cmQ;                                           // channel-major tables, interleaved by a second kernel
    if ((rc = mkid_scratch(ctx, SCR_AUX0, (size_t)batch * n_lut * 2, (void **)&cmI))) return rc;
    if ((rc = mkid_scratch(ctx, SCR_AUX1, (size_t)batch * n_lut * 2, (void **)&cmQ))) return rc;
    int cpc = 1;                                                  // channels per CTA (experiment switch; no gain measured)
    if (const char *e = getenv("MKID_DDS_CPC")) { const int v = atoi(e); if (v == 1 || v == 2 || v == 4 || v == 8) cpc = v; }
    dds_lut_kernel<<<dim3(256 / cpc, batch), 256, smem, ctx->stream>>>(d_res, d_ph, fs2, size, offset, d_tab, cpc, cmI, cmQ, d_sc);
    MKID_CHECK_LAUNCH(ctx);
    const int half = size / 2;
    dds_interleave_kernel<<<dim3(8 * ((half + 31) / 32), batch), 256, 0, ctx->stream>>>(
        (const uint32_t *)cmI, (const uint32_t *)cmQ, half, ch_shift, (uint32_t *)dI, (uint32_t *)dQ);
    MKID_CHECK_LAUNCH(ctx);
    if (scales_out) {
        MKID_CUDA(ctx, cudaMemcpyAsync(scales_out, d_sc, (size_t)batch * 256 * 8, cudaMemcpyDefault, ctx->stream));
        MKID_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    if ((rc = mkid_stage_out_finish(ctx, I_dds, (size_t)batch * n_lut * 2, dI))) return rc;
    return mkid_stage_out_finish(ctx, Q_dds, (size_t)batch * n_lut * 2, dQ);
}

extern "C" int mkid_pack_dram(mkid_ctx *ctx, const int16_t *I_dac, const int16_t *Q_dac, const int16_t *I_dds,
                              const int16_t *Q_dds, int64_t n, uint8_t *out) {
    if (!ctx) return MKID_EINVAL;
    MKID_REQUIRE(ctx, I_dac && Q_dac && I_dds && Q_dds && out && n > 0 && n % 2 == 0, "pack_dram: bad argument");
    MKID_CUDA(ctx, cudaSetDevice(ctx->device));
    const void *a, *b, *c, *d; void *o; int rc;
    if ((rc = mkid_stage_in(ctx, I_dac, n * 2, SCR_IN, &a))) return rc;
    if ((rc = mkid_stage_in(ctx, Q_dac, n * 2, SCR_IN1, &b))) return rc;
    if ((rc = mkid_stage_in(ctx, I_dds, n * 2, SCR_IN2, &c))) return rc;
    if ((rc = mkid_stage_in(ctx, Q_dds, n * 2, SCR_IN3, &d))) return rc;
    if ((rc = mkid_stage_out(ctx, out, n * 8, SCR_OUT2, false, &o))) return rc;
    const int64_t pairs = n / 2;
    const int grid = (int)std::min<int64_t>((pairs + 255) / 256, (int64_t)ctx->num_sms * 8);
    pack_dram_kernel<<<grid, 256, 0, ctx->stream>>>((const int16_t *)a, (const int16_t *)b, (const int16_t *)c,
                                                    (const int16_t *)d, pairs, (uint4 *)o);
    MKID_CHECK_LAUNCH(ctx);
    return mkid_stage_out_finish(ctx, out, n * 8, o);
}

// test hook: correctly rounded sin/cos used by the exact paths
extern "C" int mkid_sincos_cr(mkid_ctx *ctx, const double *x, int64_t n, double *s, double *c) {
    if (!ctx) return MKID_EINVAL;
    MKID_REQUIRE(ctx, x && s && c && n > 0, "sincos_cr: bad argument");
    MKID_CUDA(ctx, cudaSetDevice(ctx->device));
    const void *dx; void *ds, *dc; int rc;
    if ((rc = mkid_stage_in(ctx, x, n * 8, SCR_IN, &dx))) return rc;
    if ((rc = mkid_stage_out(ctx, s, n * 8, SCR_OUT0, false, &ds))) return rc;
    if ((rc = mkid_stage_out(ctx, c, n * 8, SCR_OUT1, false, &dc))) return rc;
    sincos_cr_test_kernel<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>((const double *)dx, n, (double *)ds, (double *)dc);
    MKID_CHECK_LAUNCH(ctx);
    if ((rc = mkid_stage_out_finish(ctx, s, n * 8, ds))) return rc;
    return mkid_stage_out_finish(ctx, c, n * 8, dc);
}
