// K4: streaming software channelizer (PFB + FFT-512 + bin select + DDS mix + 26-tap FIR /2 +
//     centre subtract + atan2 -> Fix16_13 phase), and K5: pulse detection -> 64-bit photon words.
//
// The arithmetic is the model defined in oracle/channelizer.py (the firmware data plane is absent
// from the reference; see include/mkidgpu.h for the control-plane interfaces each stage follows).
//
// K4 layout: one CTA per SM = (board, chunk of output rows), 1024 threads in four roles (PFB, 2 x FFT, channel stage)
// that work on different blocks of 8 frames (hop 256 samples, 2x oversampled) at the same time; see the comment in front
// of channelize_ws_kernel.  Per ADC sample: 4 B read from HBM once, about 130 instructions, 111 FP32 lane operations:
// the kernel is bound by the FP32 pipe, the issue slots and the shared-memory pipe together (each 65-75 % busy), not by HBM.
#include <math.h>
#include <stdlib.h>

#include <algorithm>
#include <type_traits>

#include "common.cuh"

namespace {

constexpr int NCH = 256;
constexpr int NFFT = 512;
constexpr int HOP = 256;
constexpr int PTAPS = 4;
constexpr int WIN = NFFT * PTAPS;        // 2048-sample prototype window
constexpr int FIRT = 26;
constexpr int FB = 8;                    // frames per block
constexpr int RING = 32;                 // frames kept per channel for the FIR
constexpr int PRE_ROWS = 96;             // output rows recomputed in front of each call (32 baseline + 64 look-ahead)
constexpr int RES_LO = 32;               // first resolved row of the phase buffer in mkid_chan_process
constexpr int T_START = 64;              // first absolute output index that may trigger
constexpr int FFT_STRIDE = 272;          // 16 x 17 padded float2 per 256-point FFT
constexpr int CAND_ROWS = 1024;          // rows per CTA of the candidate kernel
constexpr int64_t SEC_US = 1000000;

struct ChanDev {                         // device-resident configuration + state of one mkid_chan
    int n_boards, n_lut, Ld, M, L, W, Lw;
    float *window;        // [2048]
    float2 *tw512;        // [256]  W512^k
    float2 *tw256;        // [16][16] W256^(j*q) stored [q][j]
    float fir[FIRT];      // c_k / (2047*32767): scalar operands of the packed f32x2 FMAs of the FIR
    int16_t *bins;        // [B][256]
    uint32_t *dds;        // [B][Ld][256]  I | Q << 16 (int16 pair), channel-minor; zeroed channels hold 0
    float *gain;          // [B][256] 0 (zeroed FIR) or 1 (kept for reporting; the kernel uses the zeroed DDS entries)
    float *cen_i, *cen_q; // [B][256] 8*I_c, 8*Q_c
    int32_t *thr;         // [B][256]
    uint32_t *hist;       // 2 x [B][H + 2048] input history (packed int16 I,Q; the 2048 behind it are unused), alternating per call
    int64_t *t_next;      // [B][256]
    int H;
    // board b: the half-frame hop sign of odd bins, (-1)^(bin (f + 1)), is folded into the DDS rows with even index (always,
    // unless a table holds -32768 where it would have to be negated: then the kernel applies the sign itself)
    unsigned char fold[64];
};

}  // namespace

struct mkid_chan {
    ChanDev d;
    int64_t t_consumed = 0;              // output samples (us) produced so far per board
    // scratch owned by the object
    // phase rows and candidate mask exist twice: with mkid_chan_set_pipelined the calls alternate between the two sets,
    // so that the detection of batch k (on another stream) can run under the channelizer kernel of batch k + 1
    int16_t *phase_set[2] = {nullptr, nullptr}; size_t phase_rows_set[2] = {0, 0};
    uint32_t *mask_set[2] = {nullptr, nullptr}; size_t mask_bytes_set[2] = {0, 0};
    int mask_head_rpc[2] = {0, 0};       // rows per chunk of the K4 call that left the set's mask with deferred chunk heads (else 0)
    bool alternate = false;
    struct Pending { bool valid = false; int set = 0; int64_t rows = 0, T = 0, t_abs0 = 0; } pending;
    uint32_t *acc = nullptr; size_t acc_bytes = 0;
    uint32_t *win_cnt = nullptr; size_t win_bytes = 0;    // [B][n_win] counts then offsets
    bool detect_clean = false;           // acc / win_cnt are all zero (the emit kernels clear what they consume)
    int32_t *n_words_dev = nullptr;
    int16_t *halo = nullptr; size_t halo_bytes = 0;
    uint64_t *words_dev = nullptr; size_t words_bytes = 0;
    uint32_t *in_dev = nullptr; size_t in_bytes = 0;
    bool board_set[64] = {};
    bool fir_set = false;
    float *f32_out = nullptr;            // set by mkid_chan_set_f32_phase_out (device pointer)
    static constexpr int EV_RING = 64;   // K4 start/stop event pairs of the last EV_RING process calls
    cudaEvent_t ev_k4[2 * EV_RING] = {};
    int64_t n_calls = 0;
    int64_t n_calls_edge = 0;            // parity = the edge buffer the next call reads
};

namespace {

// ------------------------------------------------------------------------------------------
// small complex helpers
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ float2 mul_mi(float2 a) { return make_float2(a.y, -a.x); }   // a * (-i)

// forward 4-point DFT in place: (a0,a1,a2,a3) -> (X0,X1,X2,X3).  The six complex additions that do not cross the real and
// imaginary parts are packed f32x2 operations (FADD2 / FFMA2 with -1: one issue slot for both halves, same roundings
// as two scalar FADDs); the two outputs that take the -i rotation stay scalar.
#ifndef K4_SCALAR_FFT
__device__ __forceinline__ float2 padd(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 psub(float2 a, float2 b) { return __ffma2_rn(b, make_float2(-1.f, -1.f), a); }   // a - b, exact product
#else
__device__ __forceinline__ float2 padd(float2 a, float2 b) { return cadd(a, b); }
__device__ __forceinline__ float2 psub(float2 a, float2 b) { return csub(a, b); }
#endif
__device__ __forceinline__ void fft4(float2 &a0, float2 &a1, float2 &a2, float2 &a3) {
    const float2 t0 = padd(a0, a2), t1 = psub(a0, a2), t2 = padd(a1, a3), d = psub(a1, a3);
    a0 = padd(t0, t2); a2 = psub(t0, t2);
    a1 = make_float2(t1.x + d.y, t1.y - d.x);          // t1 + (-i) d
    a3 = make_float2(t1.x - d.y, t1.y + d.x);          // t1 - (-i) d
}

// forward 16-point DFT; on return X[4*k1 + k2] is in v[4*k2 + k1]
__device__ __forceinline__ void fft16(float2 (&v)[16]) {
#pragma unroll
    for (int n1 = 0; n1 < 4; ++n1) fft4(v[n1], v[n1 + 4], v[n1 + 8], v[n1 + 12]);
    // twiddle v[n1 + 4*k2] *= W16^(n1*k2)
    const float c1 = 0.92387953251128674f, s1 = 0.38268343236508977f, r2 = 0.70710678118654752f;
    v[5] = cmul(v[5], make_float2(c1, -s1));                       // W16^1
    v[9] = make_float2((v[9].x + v[9].y) * r2, (v[9].y - v[9].x) * r2);   // W16^2 = (1-i)/sqrt2
    v[13] = cmul(v[13], make_float2(s1, -c1));                     // W16^3
    v[6] = make_float2((v[6].x + v[6].y) * r2, (v[6].y - v[6].x) * r2);   // W16^2
    v[10] = mul_mi(v[10]);                                         // W16^4 = -i
    v[14] = make_float2((v[14].y - v[14].x) * r2, -(v[14].x + v[14].y) * r2);  // W16^6 = (-1-i)/sqrt2
    v[7] = cmul(v[7], make_float2(s1, -c1));                       // W16^3
    v[11] = make_float2((v[11].y - v[11].x) * r2, -(v[11].x + v[11].y) * r2);  // W16^6
    v[15] = cmul(v[15], make_float2(-c1, s1));                     // W16^9
#pragma unroll
    for (int k2 = 0; k2 < 4; ++k2) fft4(v[4 * k2], v[4 * k2 + 1], v[4 * k2 + 2], v[4 * k2 + 3]);
}

// atan2 for the phase stage: branch-free, |error| < 4e-7 rad (minimax degree-7 polynomial in t^2 on
// [0,1], fast division), exact signed-zero / axis behaviour of atan2f where the model needs it.
__device__ __forceinline__ float atan2_fast(float y, float x) {
    const float ax = fabsf(x), ay = fabsf(y);
    const float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
    const float t = __fdividef(mn, fmaxf(mx, 1e-30f));       // 0 / 0 -> 0 (both zero only for a zeroed channel at the centre)
    const float s = t * t;
    float p = -4.0562719287e-03f;                 // degree 7 in t^2: 3.8e-8 rad in exact arithmetic, 1.2e-7 rad as evaluated in float32
    p = fmaf(p, s, 2.1868927929e-02f);
    p = fmaf(p, s, -5.5920632268e-02f);
    p = fmaf(p, s, 9.6427808134e-02f);
    p = fmaf(p, s, -1.3908846780e-01f);
    p = fmaf(p, s, 1.9946606613e-01f);
    p = fmaf(p, s, -3.3329864130e-01f);
    p = fmaf(p, s, 9.9999933635e-01f);
    p = p * t;
    p = ay > ax ? 1.5707963267948966f - p : p;
    p = x < 0.f ? 3.14159265358979f - p : p;
    return copysignf(p, y);
}

// ------------------------------------------------------------------------------------------
// K4, warp-specialised form (the default).  One CTA per SM, 1024 threads in four roles that work on DIFFERENT blocks of
// 8 frames at the same time, so that the shared-memory-bound FFT and the FP32-bound PFB / channel stages overlap
// instead of alternating behind block-wide barriers (setmaxnreg gives every role the registers it needs):
//   FFT0 / FFT1  threads 0..255 / 256..511 (64 regs): group g transforms the blocks kb = g (mod 2) in place: 16
//         FFT-256 per block, 16 lanes each, two radix-16 passes; a half-warp needs only __syncwarp.
//   PFB   threads 512..767 (40 regs): thread k = branches k, k+256: sliding register window over the ADC blocks that
//         arrive by 1-D TMA bulk copies, first radix-2 stage, 16 STS.64 into exchange buffer k mod 4.
//   CHAN  threads 768..1023 (88 regs): thread c = channel c: gather its bin from exchange buffer kb mod 4, DDS mix,
//         transposed-form 26-tap FIR, centre, atan2, Fix16_13 store, candidate mask.  The role with the longest
//         dependent chains has the highest warp ids (the issue arbiter prefers them).
// Hand-over by mbarriers (256 arrivals each): u_full[buf] PFB -> FFT, x_done[buf] FFT -> CHAN, u_free[buf] CHAN -> PFB.
// No role has a barrier of its own: an ADC / DDS stage (6 / 5 of them) is refilled by an elected thread once the mbarrier
// chain u_full -> x_done -> u_free proves that every thread of the role has consumed it (see the comments at arm_adc / arm_dds).
// Every chunk is a regular chunk: the samples in front of the call come from `edge` = [input history | first 2048
// samples of the call], which is all zeros in front of the start of the stream (frames before time 0 contribute +-0).
#ifndef K4_REGS_FFT
#define K4_REGS_FFT 64
#endif
#ifndef K4_REGS_PFB
#define K4_REGS_PFB 40
#endif
#ifndef K4_REGS_CHAN
#define K4_REGS_CHAN 88
#endif
// Experiment switch -DK4_CHAN_MIX=0: the FFT group that transformed a block also gathers every channel's bin from it, mixes
// it with the DDS value (thread = channel) and leaves y[frame][channel] in the block's exchange buffer, so that the channel
// role starts from eight coalesced loads.  Output bit-identical, 70 instructions less in the channel role per block -- and
// 1.25 -> 1.49 ms: the two group-wide barriers and the longer FFT stage stretch the time a block holds its exchange
// buffer beyond what 4 buffers cover (the pipeline is bound by the latency of a block through the roles, not by the
// instruction count of the busiest role).  Default: gather and mix in the channel role.
#ifndef K4_CHAN_MIX
#define K4_CHAN_MIX 1
#endif
#define K4_STR2(x) #x
#define K4_STR(x) K4_STR2(x)
constexpr int WS_THREADS = 1024;
// Experiment switch -DK4_DDS_LDG=1: the channel role reads its DDS words straight from the L2-resident table (8 coalesced
// 4-byte loads per block and thread, prefetched one block ahead in registers) instead of through TMA stages, so that the
// 40 KiB of stages can become a FIFTH exchange buffer.  Measured (bit-identical output): the loads alone cost 1.25 -> 1.37 ms
// at 4 buffers, with 5 buffers (and the ADC refill only one block ahead) 1.41 ms.  The number of exchange buffers does
// matter -- 3 buffers: 1.35 ms, 4 buffers: 1.25 ms -- but a fifth one does not fit next to the TMA stages (226 of 227 KiB).
#ifndef K4_DDS_LDG
#define K4_DDS_LDG 0
#endif
#ifndef K4_NBUF
#define K4_NBUF (K4_DDS_LDG ? 5 : 4)
#endif
#ifndef K4_ADC_STAGES
#define K4_ADC_STAGES 6
#endif
#if K4_DDS_LDG
constexpr int WS_NBUF = K4_NBUF;
constexpr int WS_ADC_STAGES = K4_ADC_STAGES, WS_DDS_STAGES = 0;       // 8 KiB each; with the exchange buffers 220 KiB of shared memory
#else
#ifndef K4_DDS_STAGES
#define K4_DDS_STAGES 5
#endif
constexpr int WS_NBUF = K4_NBUF;
constexpr int WS_ADC_STAGES = K4_ADC_STAGES, WS_DDS_STAGES = K4_DDS_STAGES;       // 8 KiB each; with the exchange buffers 227 KiB of shared memory
static_assert(K4_DDS_STAGES == K4_NBUF + 1, "the DDS stage of block kb + 1 is refilled at x_done(kb): NBUF + 1 stages");
#endif
constexpr int WS_ADC_AHEAD = WS_ADC_STAGES - WS_NBUF;     // blocks the ADC stage refill runs ahead of the PFB role
static_assert(WS_ADC_AHEAD >= 1 && 3 * WS_NBUF + 1 <= 16, "stage / barrier budget");
#if !K4_CHAN_MIX && K4_DDS_LDG
#error "the K4_CHAN_MIX=0 experiment needs the DDS stages (-DK4_DDS_LDG=0)"
#endif

struct WsParams {
    ChanDev d;
    const uint32_t *in;      // [B][n] packed samples of this call
    const uint32_t *edge;    // [B][H + 2048]: the last H samples in front of this call (left there by the previous one)
    uint32_t *edge_next;     // the same for the NEXT call: the PFB role of every chunk copies its slice of this call's last H samples
    int64_t n;
    int64_t f0_abs;          // absolute frame index of local frame 0 (always even)
    int16_t *phase;          // [B][rows][256]
    float *phase_f32;        // optional (tests): unquantised phase of the new outputs [B][n/512][256]
    int64_t rows;            // PRE_ROWS + n/512
    int rows_per_chunk;      // multiple of 32
    int chunks_per_board;
    uint32_t *mask;          // [B][ceil(rows/32)][256] candidate bits, or nullptr
    int16_t *halo;           // [B][chunks][32][256]
    int head_deferred;       // chunks > 0 start at their own first row: the candidate bits of their first M rows are left
                             // to mask_head_kernel (which finds the baseline rows in the neighbour chunk's output)
};

#ifdef K4_WARP_ARRIVE
// One arrival per WARP (barrier counts are warps): the lanes' shared-memory accesses are ordered before the elected lane's
// releasing arrive by __syncwarp.
constexpr int WS_ARRIVALS = NCH / 32;
__device__ __forceinline__ void mk_mbar_arrive(uint64_t *bar) {
    __syncwarp();
    if ((threadIdx.x & 31) == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(mk_smem_u32(bar)) : "memory");
}
#else
constexpr int WS_ARRIVALS = NCH;
__device__ __forceinline__ void mk_mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(mk_smem_u32(bar)) : "memory");
}
#endif
// wait with a long suspend hint: the warp sleeps in hardware until the phase completes instead of polling
__device__ __forceinline__ void mk_mbar_wait_sleep(uint64_t *bar, uint32_t parity) {
#ifndef K4_SLEEP_WAIT
    mk_mbar_wait(bar, parity);
    return;
#endif
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(mk_smem_u32(bar)), "r"(parity), "r"(1000000u) : "memory");
}

// Hand-over between the roles.  Default: hardware named barriers (bar.arrive by the 256 producer threads, bar.sync by the
// 256 consumer threads, ids 1..12): a waiting warp is descheduled by the barrier unit and costs no issue slots, where the
// mbarrier polling loops (try_wait / branch / yield) took a third of all executed instructions.  -DK4_MBAR_HANDOVER keeps
// the mbarrier form.  The TMA stages (adc_full / dds_full) are transaction barriers and stay mbarriers.
#ifndef K4_MBAR_HANDOVER
struct HandOver { int id; };
__device__ __forceinline__ void ho_arrive(HandOver h) { asm volatile("bar.arrive %0, %1;" ::"r"(h.id), "r"(2 * NCH) : "memory"); }
__device__ __forceinline__ void ho_wait(HandOver h, uint32_t) { asm volatile("bar.sync %0, %1;" ::"r"(h.id), "r"(2 * NCH) : "memory"); }
#define HO_U_FULL(buf) HandOver{1 + (buf)}
#define HO_X_DONE(buf) HandOver{1 + WS_NBUF + (buf)}
#define HO_U_FREE(buf) HandOver{1 + 2 * WS_NBUF + (buf)}
#else
typedef uint64_t *HandOver;
__device__ __forceinline__ void ho_arrive(HandOver h) { mk_mbar_arrive(h); }
__device__ __forceinline__ void ho_wait(HandOver h, uint32_t parity) { mk_mbar_wait_sleep(h, parity); }
#define HO_U_FULL(buf) (&u_full[buf])
#define HO_X_DONE(buf) (&x_done[buf])
#define HO_U_FREE(buf) (&u_free[buf])
#endif

// FOLD: every board's DDS rows carry the hop sign of odd bins (ChanDev::fold): the channel role has no sign code at all
template <bool F32, bool FOLD>
__global__ void __launch_bounds__(WS_THREADS, 1) channelize_ws_kernel(WsParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2 *s_u = reinterpret_cast<float2 *>(smem_raw);                           // [4][16][FFT_STRIDE]
    float2 *s_tw = s_u + WS_NBUF * 16 * FFT_STRIDE;                               // [16][16]
    uint32_t *s_adc = reinterpret_cast<uint32_t *>(s_tw + 256);                   // [6][8][256]
    uint32_t *s_dds = s_adc + WS_ADC_STAGES * FB * NCH;                           // [5][8][256]
    // adc_full[6], dds_full[5], u_full[4], x_done[4], u_free[4]
    __shared__ __align__(8) uint64_t s_bar[WS_ADC_STAGES + WS_DDS_STAGES + 3 * WS_NBUF];
    uint64_t *adc_full = s_bar, *dds_full = s_bar + WS_ADC_STAGES, *u_full = dds_full + WS_DDS_STAGES;
    uint64_t *x_done = u_full + WS_NBUF, *u_free = x_done + WS_NBUF;
    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int i = 0; i < WS_ADC_STAGES; ++i) mk_mbar_init(&adc_full[i], 1);
        for (int i = 0; i < WS_DDS_STAGES; ++i) mk_mbar_init(&dds_full[i], 1);
        for (int i = 0; i < WS_NBUF; ++i) { mk_mbar_init(&u_full[i], WS_ARRIVALS); mk_mbar_init(&x_done[i], WS_ARRIVALS); mk_mbar_init(&u_free[i], WS_ARRIVALS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < 256) s_tw[tid] = p.d.tw256[tid];
    __syncthreads();

    const int board = blockIdx.y;
    const ChanDev &d = p.d;
    const int M = d.M;
    const int64_t row0 = (int64_t)blockIdx.x * p.rows_per_chunk;
    const int64_t row1 = min(row0 + (int64_t)p.rows_per_chunk, p.rows);
    // rows [row0,row1) are stored; rows from r_start on are computed (the M rows in front of the chunk feed the
    // rolling baseline of the trigger); row r is local output t = r - PRE_ROWS
    const bool own_head = p.mask && !(p.head_deferred && blockIdx.x > 0);       // this chunk recomputes its M baseline rows
    const int64_t r_start = own_head ? (row0 - M > 0 ? row0 - M : 0) : row0;
    const int64_t tl0 = r_start - PRE_ROWS, tl1 = row1 - PRE_ROWS;
    // first frame block: floor to a multiple of 32 frames, so that block 0 is in ring phase 0 of the FIR accumulators
    const int64_t fb_first = ((2 * tl0 - 24) >> 5) << 5;
    const int n_blocks = row0 < row1 ? (int)((2 * tl1 - fb_first + FB - 1) / FB) : 0;
    auto unpack = [](uint32_t v) -> float2 {
        return make_float2((float)(int16_t)(v & 0xFFFF), (float)(int16_t)(v >> 16));
    };

    if (tid < 2 * NCH) {
        // =============================== FFT groups ===============================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 " K4_STR(K4_REGS_FFT) ";");
        const int g = tid >> 8, lt = tid & 255, j = lt & 15;
#if !K4_CHAN_MIX
        // gather / mix stage of this group's blocks: thread lt = channel lt
        const int bin_c = d.bins[board * NCH + lt];
        const int par_c = bin_c & 1;
        const int zoff_c = par_c * FFT_STRIDE + (bin_c >> 1);
        const int ld_mask_f = d.Ld - 1;
        const uint32_t *dds_blk_f = d.dds + (size_t)board * d.Ld * NCH;
        const int dds_row0_f = (int)((p.f0_abs + fb_first) & ld_mask_f);
        auto arm_dds_f = [&](int kb) {                            // elected thread: block kb -> stage kb mod 5
            const int st = kb % WS_DDS_STAGES;
            const int row = (dds_row0_f + kb * FB) & ld_mask_f;
            mk_mbar_expect_tx(&dds_full[st], FB * NCH * 4);
            mk_bulk_g2s(s_dds + st * FB * NCH, dds_blk_f + (size_t)row * NCH, FB * NCH * 4, &dds_full[st]);
        };
        if (tid == 0)
            for (int b0 = 0; b0 < WS_DDS_STAGES && b0 < n_blocks; ++b0) arm_dds_f(b0);
#endif
        for (int kb = g; kb < n_blocks; kb += 2) {
            const int buf = kb % WS_NBUF;
            ho_wait(HO_U_FULL(buf), (uint32_t)((kb / WS_NBUF) & 1));
            float2 *reg = s_u + buf * 16 * FFT_STRIDE + (lt >> 4) * FFT_STRIDE;
            float2 v[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = reg[j + 16 * i];
            fft16(v);
            __syncwarp();
            // B_j[q] = A_j[q] * W256^(j*q) -> reg[q*17 + j];  A_j[q], q = 4*k1+k2, sits in v[4*k2+k1]
#pragma unroll
            for (int k1 = 0; k1 < 4; ++k1)
#pragma unroll
                for (int k2 = 0; k2 < 4; ++k2) {
                    const int q = 4 * k1 + k2;
                    float2 x = v[4 * k2 + k1];
                    if (q != 0) x = cmul(x, s_tw[q * 16 + j]);
                    reg[q * 17 + j] = x;
                }
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = reg[j * 17 + i];
            fft16(v);
            __syncwarp();
            // X[q + 16*r], r = 4*k1+k2 in v[4*k2+k1]
#pragma unroll
            for (int k1 = 0; k1 < 4; ++k1)
#pragma unroll
                for (int k2 = 0; k2 < 4; ++k2) reg[j + 16 * (4 * k1 + k2)] = v[4 * k2 + k1];
#if !K4_CHAN_MIX
            // every spectrum of the block is in place: gather this channel's bin of the 8 frames, remove the half-frame hop
            // phase of odd bins ((-1)^(bin*(f_abs+1)); f_abs of frame i has the parity of i), mix with conj(dds)
            asm volatile("bar.sync %0, %1;" ::"r"(13 + g), "r"(NCH) : "memory");
            {
                const float2 *zsrc = s_u + buf * 16 * FFT_STRIDE + zoff_c;
                float2 z[FB], y[FB];
#pragma unroll
                for (int i = 0; i < FB; ++i) z[i] = zsrc[(2 * i) * FFT_STRIDE];
                const int st = kb % WS_DDS_STAGES;
                mk_mbar_wait_sleep(&dds_full[st], (uint32_t)((kb / WS_DDS_STAGES) & 1));
                const uint32_t *dds_c = s_dds + st * FB * NCH + lt;
#pragma unroll
                for (int i = 0; i < FB; ++i) {
                    const float2 dvi = unpack(dds_c[i * NCH]);      // zeroed channels hold (0, 0): y = +-0, w = +0 as in the model
                    float2 zz = z[i];
                    if ((i & 1) == 0 && par_c && !d.fold[board]) { zz.x = -zz.x; zz.y = -zz.y; }     // even i: f_abs + 1 odd
                    y[i].x = zz.x * dvi.x + zz.y * dvi.y;
                    y[i].y = zz.y * dvi.x - zz.x * dvi.y;
                }
                // all gathers of the group are done (and its DDS values read): the mixed samples replace the spectra,
                // y[frame][channel] at the start of the buffer; the DDS stage is re-armed five blocks ahead
                asm volatile("bar.sync %0, %1;" ::"r"(13 + g), "r"(NCH) : "memory");
                if (lt == 0 && kb + WS_DDS_STAGES < n_blocks) arm_dds_f(kb + WS_DDS_STAGES);
                float2 *ydst = s_u + buf * 16 * FFT_STRIDE + lt;
#pragma unroll
                for (int i = 0; i < FB; ++i) ydst[i * NCH] = y[i];
            }
#endif
            ho_arrive(HO_X_DONE(buf));
        }
        return;
    }

    if (tid < 3 * NCH) {
        // =============================== PFB: polyphase filter + first radix-2 stage ===============================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 " K4_STR(K4_REGS_PFB) ";");
        const int k = tid - 2 * NCH;                                  // branches k and k + 256
        // the input history of the NEXT call (its last H samples), a slice per chunk, once this role has nothing else to do:
        // it was a kernel of its own between two channelizer launches (6.6 us of a 1.25 ms step)
        auto save_history = [&]() {
            const int n16 = d.H / 4, per = (n16 + (int)gridDim.x - 1) / (int)gridDim.x;
            const uint4 *src = reinterpret_cast<const uint4 *>(p.in + (size_t)board * p.n + (p.n - d.H));
            uint4 *dst = reinterpret_cast<uint4 *>(p.edge_next + (size_t)board * (d.H + 2048));
            const int i1 = min(n16, ((int)blockIdx.x + 1) * per);
            for (int i = (int)blockIdx.x * per + k; i < i1; i += NCH) dst[i] = src[i];
        };
        if (n_blocks == 0) { save_history(); return; }
        float hA[PTAPS], hB[PTAPS];
#pragma unroll
        for (int q = 0; q < PTAPS; ++q) { hA[q] = d.window[NFFT * q + k]; hB[q] = d.window[NFFT * q + HOP + k]; }
        const float2 w512 = d.tw512[k];
        const uint32_t *in = p.in + (size_t)board * p.n;
        const uint32_t *edge = p.edge + (size_t)board * (d.H + 2048) + d.H;   // edge[idx] valid for -H <= idx < 0
        auto arm_adc = [&](int blk) {                                 // elected thread: block blk -> stage blk mod 6
            const int st = blk % WS_ADC_STAGES;
            const int64_t s0 = HOP * (fb_first + (int64_t)blk * FB);
            const uint32_t *src = s0 >= 0 ? in + s0 : edge + s0;
            mk_mbar_expect_tx(&adc_full[st], FB * HOP * 4);
            mk_bulk_g2s(s_adc + st * FB * NCH, src, FB * HOP * 4, &adc_full[st]);
        };
        if (k == 0)
            for (int b0 = 0; b0 < WS_ADC_STAGES && b0 < n_blocks; ++b0) arm_adc(b0);
        // warm-up: s[j] = x[256*(f+1) - 2048 + k + 256*j], j = 0..6 for f = fb_first
        float2 sw[8];
#pragma unroll
        for (int j = 0; j < 7; ++j) {
            const int64_t idx = HOP * (fb_first + 1) - WIN + k + HOP * j;
            sw[j] = unpack(idx >= 0 ? in[idx] : edge[idx]);
        }
        sw[7] = make_float2(0.f, 0.f);
        for (int blk = 0; blk < n_blocks; ++blk) {
            const int buf = blk % WS_NBUF;
            if (blk >= WS_NBUF) {
                ho_wait(HO_U_FREE(buf), (uint32_t)((blk / WS_NBUF - 1) & 1));   // CHAN has gathered block blk - 4
                // u_free(blk - NBUF) follows x_done(blk - NBUF), which follows u_full(blk - NBUF): EVERY PFB thread has
                // finished block blk - NBUF, so its ADC stage can be refilled (block blk + STAGES - NBUF goes there) without
                // a barrier
                if (k == 0 && blk + WS_ADC_AHEAD < n_blocks) arm_adc(blk + WS_ADC_AHEAD);
            }
            const int st = blk % WS_ADC_STAGES;
            mk_mbar_wait_sleep(&adc_full[st], (uint32_t)((blk / WS_ADC_STAGES) & 1));
            const uint32_t *adc_c = s_adc + st * FB * NCH + k;
            float2 *ub = s_u + buf * 16 * FFT_STRIDE + k;
#pragma unroll
            for (int i = 0; i < FB; ++i) {
                sw[(i + 7) & 7] = unpack(adc_c[i * NCH]);
                float2 u0 = make_float2(0.f, 0.f), u1 = make_float2(0.f, 0.f);
#pragma unroll
                for (int q = 0; q < PTAPS; ++q) {
                    u0 = __ffma2_rn(make_float2(hA[q], hA[q]), sw[(2 * q + i) & 7], u0);
                    u1 = __ffma2_rn(make_float2(hB[q], hB[q]), sw[(2 * q + 1 + i) & 7], u1);
                }
                ub[(2 * i) * FFT_STRIDE] = cadd(u0, u1);                      // even bins
                ub[(2 * i + 1) * FFT_STRIDE] = cmul(csub(u0, u1), w512);      // odd bins
            }
            ho_arrive(HO_U_FULL(buf));
        }
        save_history();
        return;
    }

    // =============================== CHAN: thread = channel ===============================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 " K4_STR(K4_REGS_CHAN) ";");
    if (n_blocks == 0) return;
    const int c = tid - 3 * NCH;                                  // channel
    const int bin = d.bins[board * NCH + c];
    const int par = bin & 1;
    const int zoff = par * FFT_STRIDE + (bin >> 1);                // + 2*i*FFT_STRIDE per frame, + buffer base
    const bool hop_folded = d.fold[board] != 0;
    const float cen_i = d.cen_i[board * NCH + c], cen_q = d.cen_q[board * NCH + c];
    int16_t *phase = p.phase + (size_t)board * p.rows * NCH + c;
    // fused candidate mask (K5c): bit (r & 31) of mask[r >> 5][c] iff M*raw[r] - sum_{k=1..M} raw[r-k] < M*thr
    const int thM = M * d.thr[board * NCH + c];
    const int64_t r_eval0 = (p.mask && !own_head) ? row0 + M : (row0 > RES_LO ? row0 : RES_LO);
    int16_t *halo = p.halo ? p.halo + (((size_t)board * p.chunks_per_board + blockIdx.x) * 32) * NCH + c : nullptr;
    uint32_t *mk = p.mask ? p.mask + (size_t)board * ((p.rows + 31) >> 5) * NCH + c : nullptr;
    int S = 0;
    uint32_t bits = 0;
    const int ld_mask = d.Ld - 1;                               // Ld is a power of two
    const uint32_t *dds_blk = d.dds + (size_t)board * d.Ld * NCH;
    const int dds_row0 = (int)((p.f0_abs + fb_first) & ld_mask);   // f_abs mod Ld of the first frame of block 0
#if K4_DDS_LDG
    // DDS words of a block: frames kb*8 .. kb*8+7 of this channel (the rows of a block never wrap: block starts are
    // multiples of 8 frames), from the L2-resident table
    auto load_dds = [&](int kb, uint32_t (&dst)[FB]) {
        const uint32_t *src = dds_blk + (size_t)((dds_row0 + kb * FB) & ld_mask) * NCH + c;
#pragma unroll
        for (int i = 0; i < FB; ++i) dst[i] = __ldg(src + i * NCH);
    };
    uint32_t ddsA[FB], ddsB[FB];
    load_dds(0, ddsA);
#else
    auto arm_dds = [&](int kb) {                                  // elected thread: block kb -> stage kb mod 5
        const int st = kb % WS_DDS_STAGES;
        const int row = (dds_row0 + kb * FB) & ld_mask;
        mk_mbar_expect_tx(&dds_full[st], FB * NCH * 4);
        mk_bulk_g2s(s_dds + st * FB * NCH, dds_blk + (size_t)row * NCH, FB * NCH * 4, &dds_full[st]);
    };
#if K4_CHAN_MIX
    if (c == 0)
        for (int b0 = 0; b0 < WS_DDS_STAGES && b0 < n_blocks; ++b0) arm_dds(b0);
#endif
#endif
    // ---- state in chunk-relative rows (32-bit): row r = row0 + rl
    const int n_rows = (int)(row1 - row0);
    const int rl_start = (int)(r_start - row0);                       // <= 0: first computed row
    const int rl_eval0 = (int)(r_eval0 - row0);                       // first row whose trigger condition is evaluated
    int16_t *phase_c = phase + row0 * NCH;                            // this thread's column at the chunk's first row
    uint32_t *mk_c = mk ? mk + (row0 >> 5) * NCH : nullptr;           // row0 is a multiple of 32
    float *f32_c = (F32 && p.phase_f32) ? p.phase_f32 + ((size_t)board * (p.rows - PRE_ROWS) + (row0 - PRE_ROWS)) * NCH + c : nullptr;
    int rl = (int)((fb_first >> 1) + PRE_ROWS - row0);                // relative row of the first output of block 0
    const int rl_fast = rl_eval0 > M ? rl_eval0 : M;                  // from here on no boundary cases

    // Transposed-form FIR: every new frame adds its contribution to the 13 outputs it belongs to.
    // acc[] holds 16 output accumulators (4 completing in this block + 12 pending); the slot of the
    // output with block-relative index m is (m + 4*block) mod 16, static inside each of the 4 ring
    // phases RB = 8*(block mod 4).
    float2 acc[16];
#pragma unroll
    for (int m = 0; m < 16; ++m) acc[m] = make_float2(0.f, 0.f);

    auto channel_stage = [&](auto RBc, const float2 *xbuf, auto dds_at, HandOver free_bar) {
        constexpr int RB = decltype(RBc)::value;
        constexpr int A0 = RB / 2;                                       // slot of output m = 0
        // gather the bin, remove the half-frame hop phase of odd bins ((-1)^(bin*(f_abs+1)); f_abs of
        // frame i has the parity of i because block starts and call starts are even), mix with conj(dds)
        float2 y[FB];
#if K4_CHAN_MIX
        const float2 *zsrc = xbuf + zoff;
        float2 z[FB];
#pragma unroll
        for (int i = 0; i < FB; ++i) z[i] = zsrc[(2 * i) * FFT_STRIDE];
        if (FOLD) {                                             // (compile time) the sign of odd bins is in the DDS rows
#pragma unroll
            for (int i = 0; i < FB; ++i) {
                const float2 dvi = unpack(dds_at(i));           // zeroed channels hold (0, 0): y = +-0, w = +0 as in the model
                y[i].x = z[i].x * dvi.x + z[i].y * dvi.y;
                y[i].y = z[i].y * dvi.x - z[i].x * dvi.y;
            }
        } else {
#pragma unroll
            for (int i = 0; i < FB; ++i) {
                const float2 dvi = unpack(dds_at(i));
                float2 zz = z[i];
                if ((i & 1) == 0 && par && !hop_folded) { zz.x = -zz.x; zz.y = -zz.y; }         // even i: f_abs + 1 odd
                y[i].x = zz.x * dvi.x + zz.y * dvi.y;
                y[i].y = zz.y * dvi.x - zz.x * dvi.y;
            }
        }
#else
        // the FFT group has left the mixed samples y[frame][channel] in the buffer: coalesced, conflict-free loads
#pragma unroll
        for (int i = 0; i < FB; ++i) y[i] = xbuf[i * NCH + c];
#endif
        ho_arrive(free_bar);                                      // the bins of this block are in registers: the buffer may be refilled
        // output t = fb/2 + m uses frames 2t+1-25+k, k = 0..25: frame fb+i carries tap k = i - 2m + 24
#pragma unroll
        for (int k = 0; k < FIRT; ++k) {
#pragma unroll
            for (int m = 0; m < 16; ++m) {
                const int i = k - 24 + 2 * m;
                if (i >= 0 && i < FB) acc[(A0 + m) & 15] = __ffma2_rn(make_float2(d.fir[k], d.fir[k]), y[i], acc[(A0 + m) & 15]);
            }
        }
        float ar[4], ai[4];
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
            ar[jj] = acc[(A0 + jj) & 15].x; ai[jj] = acc[(A0 + jj) & 15].y;
            acc[(A0 + jj) & 15] = make_float2(0.f, 0.f);                  // becomes output m = 12 + jj of the next block
        }
        if (rl + 3 < rl_start || rl >= n_rows) return;
        const bool fast = rl >= rl_fast && rl + 3 < n_rows;
        // rows leaving the rolling baseline while these 4 outputs enter it (M >= 4, so they were stored by
        // an earlier block)
        int old[4] = {0, 0, 0, 0};
        if (mk) {
            if (fast) {
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) old[jj] = phase_c[(rl + jj - M) * NCH];
            } else {
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    const int ro = rl + jj - M;
                    if (ro >= rl_start && rl + jj >= rl_start) old[jj] = ro >= 0 ? phase_c[ro * NCH] : halo[(32 + ro) * NCH];
                }
            }
        }
        int raw[4];
        float ph[4];
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
            const float a = ar[jj] - cen_i, b = ai[jj] - cen_q;
            ph[jj] = atan2_fast(b, a);
            raw[jj] = __float2int_rn(ph[jj] * 8192.0f);
        }
        if (fast) {
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
                phase_c[(rl + jj) * NCH] = (int16_t)raw[jj];
                if (F32) { if (f32_c && rl + jj + (int)(row0 - PRE_ROWS) >= 0) f32_c[(size_t)(rl + jj) * NCH] = ph[jj]; }
                if (mk) {
                    if ((M * raw[jj] - S) < thM) bits |= 1u << ((rl + jj) & 31);
                    S += raw[jj] - old[jj];
                }
            }
            if (mk && (rl & 31) == 28) { mk_c[(rl >> 5) * NCH] = bits; bits = 0; }
        } else {
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
                const int r = rl + jj;
                if (r < rl_start || r >= n_rows) continue;
                if (r >= 0) phase_c[r * NCH] = (int16_t)raw[jj];
                else halo[(32 + r) * NCH] = (int16_t)raw[jj];
                if (F32) { if (f32_c && r >= 0 && r + (int)(row0 - PRE_ROWS) >= 0) f32_c[(size_t)r * NCH] = ph[jj]; }
                if (mk) {
                    if (r >= rl_eval0 && (M * raw[jj] - S) < thM) bits |= 1u << (r & 31);
                    S += raw[jj] - old[jj];
                    if (r >= 0 && ((r & 31) == 31 || r == n_rows - 1)) { mk_c[(r >> 5) * NCH] = bits; bits = 0; }
                }
            }
        }
    };

    // The ring phase of block kb is 8 * (kb mod 4) (fb_first is a multiple of 32 frames): four statically addressed
    // copies in sequence, so that the accumulators never move between registers.
#if K4_DDS_LDG
    auto one_block = [&](auto RBc, int kb, uint32_t (&cur)[FB], uint32_t (&nxt)[FB]) {
        const int buf = kb % WS_NBUF;
        if (kb + 1 < n_blocks) load_dds(kb + 1, nxt);            // in flight while this block is awaited and gathered
        ho_wait(HO_X_DONE(buf), (uint32_t)((kb / WS_NBUF) & 1));
        channel_stage(RBc, s_u + buf * 16 * FFT_STRIDE, [&](int i) { return cur[i]; }, HO_U_FREE(buf));
        rl += 4;
    };
    for (int kb = 0; kb < n_blocks; kb += 4) {
        one_block(std::integral_constant<int, 0>{}, kb, ddsA, ddsB);
        if (kb + 1 >= n_blocks) break;
        one_block(std::integral_constant<int, 8>{}, kb + 1, ddsB, ddsA);
        if (kb + 2 >= n_blocks) break;
        one_block(std::integral_constant<int, 16>{}, kb + 2, ddsA, ddsB);
        if (kb + 3 >= n_blocks) break;
        one_block(std::integral_constant<int, 24>{}, kb + 3, ddsB, ddsA);
    }
#else
    auto one_block = [&](auto RBc, int kb) {
        const int buf = kb % WS_NBUF;
        ho_wait(HO_X_DONE(buf), (uint32_t)((kb / WS_NBUF) & 1));
        // x_done(kb) follows u_full(kb), which follows u_free(kb - 4): EVERY CHAN thread has read the DDS values of block
        // kb - 4, so that stage can be refilled (5 stages: block kb + 1 goes there) without a barrier
        const int st = kb % WS_DDS_STAGES;
#if K4_CHAN_MIX
        if (c == 0 && kb >= WS_NBUF && kb + 1 < n_blocks) arm_dds(kb + 1);
        mk_mbar_wait_sleep(&dds_full[st], (uint32_t)((kb / WS_DDS_STAGES) & 1));
#endif
        const uint32_t *dds_c = s_dds + st * FB * NCH + c;
        channel_stage(RBc, s_u + buf * 16 * FFT_STRIDE, [&](int i) { return dds_c[i * NCH]; }, HO_U_FREE(buf));
        rl += 4;
    };
    for (int kb = 0; kb < n_blocks; kb += 4) {
        one_block(std::integral_constant<int, 0>{}, kb);
        if (kb + 1 >= n_blocks) break;
        one_block(std::integral_constant<int, 8>{}, kb + 1);
        if (kb + 2 >= n_blocks) break;
        one_block(std::integral_constant<int, 16>{}, kb + 2);
        if (kb + 3 >= n_blocks) break;
        one_block(std::integral_constant<int, 24>{}, kb + 3);
    }
#endif
}

// edge[b] = [history (H samples) | 2048 unused].  Two edge buffers alternate from call to call: call k reads buffer k & 1 (the H
// samples in front of it) and its PFB roles leave the last H samples of its own input in the other one.  Blocks of 8 frames
// start at multiples of 2048 samples, so no TMA copy straddles the start of the call.  (Round 1 copied history after K4 and
// the head of the call in front of it; later one edge_prep kernel in front of K4 did both; now nothing stands between two
// channelizer launches.)

// Candidate bits of the first M rows of every chunk > 0 of a channelizer call (head_deferred): their rolling baseline is the
// last M rows of the NEIGHBOUR chunk, which that chunk's CTA was still computing while K4 ran.  K4 used to recompute those M
// rows itself in front of every chunk (M of 456 + M rows with one board per GPU); now it leaves the low M bits of the
// chunk's first mask word clear and this kernel fills them from the stored rows - the same integer arithmetic on the same
// int16 values.  Thread = channel, block = (chunk, board); runs in front of the resolver.
__global__ void __launch_bounds__(NCH) mask_head_kernel(const int16_t *__restrict__ phase, int64_t rows, int rpc, int M,
                                                        const int32_t *__restrict__ thr, uint32_t *mask) {
    const int board = blockIdx.y, c = threadIdx.x;
    const int64_t row0 = (int64_t)(blockIdx.x + 1) * rpc;
    if (row0 >= rows) return;
    const int16_t *ph = phase + (size_t)board * rows * NCH + c;
    const int thM = M * thr[board * NCH + c];
    int S = 0;
    for (int k = 1; k <= M; ++k) S += ph[(row0 - k) * NCH];
    const int n = (int)min((int64_t)M, rows - row0);
    uint32_t bits = 0;
    for (int i = 0; i < n; ++i) {
        const int raw = ph[(row0 + i) * NCH];
        if (M * raw - S < thM) bits |= 1u << i;
        S += raw - ph[(row0 + i - M) * NCH];
    }
    uint32_t *w = mask + ((size_t)board * ((rows + 31) >> 5) + (row0 >> 5)) * NCH + c;
    *w = (M >= 32 ? 0u : *w & ~((1u << M) - 1u)) | bits;
}

// ------------------------------------------------------------------------------------------
// K5c: candidate mask from the phase rows.  bit (r & 31) of mask[b][r >> 5][c] is set iff
//      M*raw[r] - sum_{k=1..M} raw[r-k] < M*thr[c]   (rows r >= M)
__global__ void __launch_bounds__(256) candidates_kernel(const int16_t *__restrict__ phase, int64_t rows, int M,
                                                         const int32_t *__restrict__ thr, uint32_t *mask) {
    const int c = threadIdx.x, board = blockIdx.y;
    const int64_t r0 = (int64_t)blockIdx.x * CAND_ROWS;
    const int64_t r1 = min(r0 + CAND_ROWS, rows);
    const int16_t *ph = phase + (size_t)board * rows * NCH;
    const int64_t n_groups = (rows + 31) >> 5;
    uint32_t *mk = mask + (size_t)board * n_groups * NCH;
    const int th = M * thr[board * NCH + c];
    int S = 0;
    for (int k = 1; k <= M; ++k) {
        const int64_t r = r0 - k;
        if (r >= 0) S += ph[r * NCH + c];
    }
    for (int64_t rg = r0; rg < r1; rg += 32) {
        uint32_t bits = 0;
#pragma unroll 8
        for (int b = 0; b < 32; ++b) {
            const int64_t r = rg + b;
            if (r < r1) {
                const int v = ph[r * NCH + c];
                if (r >= M && (M * v - S) < th) bits |= 1u << b;
                S += v;
                if (r >= M) S -= ph[(r - M) * NCH + c];
            }
        }
        mk[(rg >> 5) * NCH + c] = bits;
    }
}

// K5a: greedy hold-off per channel (sequential in time per channel).  One CTA per 8 channels (256 CTAs for 8 boards):
// all 8 warps stream the CTA's 32-byte wide column of the candidate mask through a double-buffered shared-memory
// tile of 512 mask words (cp.async, 16 B per copy; a batch of 2^16 rows is 4 tiles, so the walk costs 4 memory round
// trips) and reduce it to one "word is non-zero" bit per (channel, word).  Lane 0 of warp w then walks the non-zero
// words of channel w only: the per-trigger chain "next candidate at or after the end of the dead time" runs at
// shared-memory latency and skips the dead time in one step.
// Two shapes.  RES_TW = 512 / 256 threads (a warp per channel, 4 tiles per 2^16 rows: the shortest latency when the kernel
// has the GPU to itself) and RES_TW = 128 / 64 threads (four resolver lanes per warp, 8 KiB of tiles: 28 CTAs per SM instead
// of 7, the better throughput when the kernel is confined to the few SMs the channelizer kernel of the next batch leaves).
constexpr int RES_CH = 8;                             // channels per CTA
template <int RES_TW>                                 // mask words (x 32 rows) per tile; RES_TW / 2 threads
__global__ void __launch_bounds__(RES_TW / 2) resolve_kernel(const uint32_t *__restrict__ mask, int64_t rows, int64_t r_lo,
                                                      int64_t r_hi, int64_t t_abs0, int L, int Lw, int n_win,
                                                      int64_t *t_next, uint32_t *acc, uint32_t *win_cnt) {
    __shared__ __align__(16) uint32_t tile[2][RES_TW][RES_CH];
    __shared__ __align__(16) uint16_t summ[RES_CH][RES_TW / 16];   // per channel: non-zero flags of the current tile
    const int board = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int c0 = blockIdx.x * RES_CH;
    const int64_t n_groups = (rows + 31) >> 5;
    const uint32_t *mk = mask + (size_t)board * n_groups * NCH + c0;
    uint32_t *ac = acc + (size_t)board * n_win * NCH;
    uint32_t *wc = win_cnt + (size_t)board * (n_win + 1);
    const int g_lo = (int)(r_lo >> 5), g_hi = (int)((r_hi + 31) >> 5);
    const int n_tiles = (g_hi - g_lo + RES_TW - 1) / RES_TW;
    auto prefetch = [&](int k) {                       // tile k -> buffer k & 1: RES_TW rows of 2 x 16 B
        const int g0 = g_lo + k * RES_TW;
        for (int i = tid; i < RES_TW * 2; i += RES_TW / 2) {
            const int u = i >> 1, q = i & 1;
            if (g0 + u < g_hi) {
                const uint32_t dst = mk_smem_u32(&tile[k & 1][u][q * 4]);
                const uint32_t *src = mk + (size_t)(g0 + u) * NCH + q * 4;
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
            } else {
                *reinterpret_cast<uint4 *>(&tile[k & 1][u][q * 4]) = make_uint4(0, 0, 0, 0);
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    // resolver lanes: 8 per CTA, spread evenly over its warps (lane 0 of each of 8 warps, or lanes 0, 8, 16, 24 of 2 warps);
    // each owns channel c0 + cl.  Rows are 32-bit inside a call.
    constexpr int RES_WARPS = RES_TW / 64, RES_STEP = 4 * RES_WARPS;
    const bool resolver = (lane % RES_STEP) == 0;
    const int cl = warp * (RES_CH / RES_WARPS) + lane / RES_STEP;      // channel inside the CTA (resolver lanes)
    const int c = c0 + cl;
    const int rlo = (int)r_lo, rhi = (int)r_hi;
    int64_t tn = 0;
    int lo = rlo;                                      // first row this channel may trigger on
    int last_r = -1, wi = 0, w_end = rlo + Lw;         // last trigger of this call, window of the next one
    auto first_allowed = [&]() -> int {
        const int64_t t_min = tn > T_START ? tn : (int64_t)T_START;   // hold-off from earlier calls, start-up guard
        const int64_t r = t_min - t_abs0;
        return r > (int64_t)rlo ? (r > (int64_t)rhi ? rhi : (int)r) : rlo;
    };
    if (resolver) { tn = t_next[board * NCH + c]; lo = first_allowed(); }
    prefetch(0);
    for (int k = 0; k < n_tiles; ++k) {
        if (k + 1 < n_tiles) prefetch(k + 1);
        else asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 1;" ::: "memory");
        __syncthreads();
        uint32_t(*t)[RES_CH] = tile[k & 1];
        {   // non-zero flags: thread = (channel, group of 16 words)
            const int ch = tid & (RES_CH - 1), grp = tid >> 3;
            uint32_t part = 0;
#pragma unroll
            for (int i = 0; i < 16; ++i) part |= (t[grp * 16 + i][ch] != 0u ? 1u : 0u) << i;
            summ[ch][grp] = (uint16_t)part;
        }
        __syncthreads();
        const int g0 = g_lo + k * RES_TW;
        if (resolver && lo < ((g0 + RES_TW) << 5)) {   // (else: the whole tile lies in this channel's dead time)
            uint32_t *sw = reinterpret_cast<uint32_t *>(&summ[cl][0]);     // RES_TW / 32 words of flags, only this lane's
            // the chain below is serial per channel (about 40 rounds per 2^16 rows): 32-bit rows, no division
            for (;;) {
                // first non-zero word that does not end before lo
                const int u_min = max((lo >> 5) - g0, 0);
                if (u_min >= RES_TW) break;
                int u = -1;
                for (int j = u_min >> 5; j < RES_TW / 32; ++j) {
                    uint32_t v = sw[j];
                    if (j == (u_min >> 5)) v &= 0xFFFFFFFFu << (u_min & 31);
                    if (v) { u = 32 * j + __ffs(v) - 1; break; }
                }
                if (u < 0) break;
                uint32_t w = t[u][cl];
                const int r_base = (g0 + u) << 5;
                if (lo > r_base) w &= 0xFFFFFFFFu << (lo - r_base);          // lo - r_base < 32 here
                if (r_base + 32 > rhi) { const int keep = rhi - r_base; w = keep <= 0 ? 0u : (w & (0xFFFFFFFFu >> (32 - keep))); }
                if (!w) { lo = max(lo, r_base + 32); continue; }
                const int r = r_base + (__ffs(w) - 1);
                last_r = r;
                while (r >= w_end) { ++wi; w_end += Lw; }                    // window of the trigger (rows only grow)
                ac[(size_t)wi * NCH + c] = (uint32_t)(r + 1);
                atomicAdd(&wc[wi], 1u);
                lo = min(r + L, rhi);                    // hold-off: L >= 32, so the next trigger is in a later word
            }
        }
        __syncthreads();
    }
    if (resolver && last_r >= 0) tn = t_abs0 + last_r + L;
    if (resolver) t_next[board * NCH + c] = tn;
}

// K5s: per board, add the end-of-second events and turn per-window counts into offsets.
__global__ void __launch_bounds__(256) scan_kernel(uint32_t *win_cnt, int n_win, int64_t r_lo, int64_t r_hi,
                                                   int64_t t_abs0, int Lw, int32_t *n_words, int64_t words_cap, int32_t *overflow) {
    __shared__ uint32_t s_part[256];
    const int board = blockIdx.x, tid = threadIdx.x;
    uint32_t *wc = win_cnt + (size_t)board * (n_win + 1);
    // second boundaries B (multiples of 1e6, B > 0) with t_abs0 + r_lo <= B < t_abs0 + r_hi
    if (tid == 0) {
        const int64_t ta = t_abs0 + r_lo, tb = t_abs0 + r_hi;
        int64_t B = ((ta + SEC_US - 1) / SEC_US) * SEC_US;
        if (ta < 0) B = 0;
        for (; B < tb; B += SEC_US)
            if (B > 0) wc[(B - ta) / Lw] += 1;
    }
    __syncthreads();
    const int per = (n_win + 255) / 256;
    uint32_t sum = 0;
    for (int i = 0; i < per; ++i) {
        const int w = tid * per + i;
        if (w < n_win) sum += wc[w];
    }
    s_part[tid] = sum;
    __syncthreads();
    if (tid == 0) {
        uint32_t run = 0;
        for (int i = 0; i < 256; ++i) { uint32_t v = s_part[i]; s_part[i] = run; run += v; }
        n_words[board] = (int32_t)run;
        if ((int64_t)run > words_cap) atomicOr(overflow, 1);          // sticky: the emit kernels drop the words beyond the capacity
    }
    __syncthreads();
    uint32_t run = s_part[tid];
    for (int i = 0; i < per; ++i) {
        const int w = tid * per + i;
        if (w < n_win) { uint32_t v = wc[w]; wc[w] = run; run += v; }
    }
}

// K5b: build the photon words of one window and write them in (time, channel) order.
__global__ void __launch_bounds__(256) emit_kernel(const int16_t *__restrict__ phase, int64_t rows,
                                                   uint32_t *acc, uint32_t *win_off,
                                                   int n_win, int64_t r_lo, int64_t r_hi, int64_t t_abs0, int M, int W,
                                                   int Lw, uint64_t *words, int64_t words_cap) {
    __shared__ uint32_t s_keys[NCH + 2];
    __shared__ int s_n;
    const int board = blockIdx.y, wi = blockIdx.x, c = threadIdx.x;
    const int16_t *ph = phase + (size_t)board * rows * NCH;
    const uint32_t a = acc[((size_t)board * n_win + wi) * NCH + c];
    if (a) acc[((size_t)board * n_win + wi) * NCH + c] = 0u;        // left clean for the next batch (no memset per call)
    const int64_t w_lo = r_lo + (int64_t)wi * Lw;          // first row of the window
    if (c == 0) s_n = 0;
    __syncthreads();
    uint32_t key = 0xFFFFFFFFu;
    uint64_t word = 0;
    if (a) {
        const int64_t r = (int64_t)a - 1;
        int S = 0;
#pragma unroll 4
        for (int k = 1; k <= M; ++k) S += ph[(r - k) * NCH + c];
        int vmin = ph[r * NCH + c];
        int64_t tp = r;
#pragma unroll 8
        for (int j = 1; j < W; ++j) {
            const int v = ph[(r + j) * NCH + c];
            if (v < vmin) { vmin = v; tp = r + j; }
        }
        const double y1 = (double)ph[(tp - 1) * NCH + c], y2 = (double)vmin, y3 = (double)ph[(tp + 1) * NCH + c];
        const double den = __dsub_rn(__dadd_rn(y3, y1), __dmul_rn(2.0, y2));
        double y4 = y2;
        if (den != 0.0) {
            const double dy = __dsub_rn(y3, y1);
            y4 = __dsub_rn(y2, __ddiv_rn(__dmul_rn(0.125, __dmul_rn(dy, dy)), den));
        }
        const int peak = (__double2int_rz(__dmul_rn(y4, 0.0625)) + 2048) & 0xFFF;
        const int p1 = (vmin / 16 + 2048) & 0xFFF;
        const int base = (S / (16 * M) + 2048) & 0xFFF;
        const int64_t t = t_abs0 + r;
        const uint32_t ts = (uint32_t)(t % SEC_US);
        word = ((uint64_t)c << 56) | ((uint64_t)peak << 44) | ((uint64_t)p1 << 32) | ((uint64_t)base << 20) | ts;
        key = ((uint32_t)(r - w_lo) << 9) | (uint32_t)(c + 1);
        s_keys[atomicAdd(&s_n, 1)] = key;
    }
    // end-of-second event inside this window?
    if (c == 0) {
        const int64_t ta = t_abs0 + w_lo;
        const int64_t tb = min(t_abs0 + w_lo + Lw, t_abs0 + r_hi);
        int64_t B = ta <= 0 ? SEC_US : ((ta + SEC_US - 1) / SEC_US) * SEC_US;
        if (B < tb) s_keys[atomicAdd(&s_n, 1)] = ((uint32_t)(B - ta) << 9);          // channel field 0: before all channels
    }
    const uint32_t off = win_off[(size_t)board * (n_win + 1) + wi];
    __syncthreads();
    const int n = s_n;
    if (c == 0) win_off[(size_t)board * (n_win + 1) + wi] = 0u;     // (every thread of the CTA has read it)
    uint64_t *out = words + (size_t)board * words_cap;
    if (a) {
        int rank = 0;
        for (int i = 0; i < n; ++i) rank += s_keys[i] < key;
        if ((int64_t)off + rank < words_cap) out[off + rank] = word;
    }
    if (c == 0) {   // the EOS word, if any
        const int64_t ta = t_abs0 + w_lo;
        const int64_t tb = min(t_abs0 + w_lo + Lw, t_abs0 + r_hi);
        int64_t B = ta <= 0 ? SEC_US : ((ta + SEC_US - 1) / SEC_US) * SEC_US;
        if (B < tb) {
            const uint32_t k = ((uint32_t)(B - ta) << 9);
            int rank = 0;
            for (int i = 0; i < n; ++i) rank += s_keys[i] < k;
            if ((int64_t)off + rank < words_cap) out[off + rank] = ~0ull;
        }
    }
}

// K5b for M <= 32 and W <= 32 (the defaults are 20 and 32): one WARP per photon word.  The per-thread version above
// walks its M + W phase samples (one 32-byte sector each, 512 B apart) a few loads at a time; here the lanes of a warp
// fetch the baseline rows and the peak window of a trigger in two instructions, two triggers per round, and reduce
// with REDUX.  Same integer / double arithmetic, same word order.
__global__ void __launch_bounds__(256) emit_warp_kernel(const int16_t *__restrict__ phase, int64_t rows,
                                                        uint32_t *acc, uint32_t *win_off,
                                                        int n_win, int64_t r_lo, int64_t r_hi, int64_t t_abs0, int M, int W,
                                                        int Lw, uint64_t *words, int64_t words_cap) {
    __shared__ uint32_t s_keys[NCH + 2];
    __shared__ int s_n;
    const int board = blockIdx.y, wi = blockIdx.x, c = threadIdx.x, lane = c & 31, warp = c >> 5;
    const int16_t *ph = phase + (size_t)board * rows * NCH;
    const uint32_t a = acc[((size_t)board * n_win + wi) * NCH + c];
    if (a) acc[((size_t)board * n_win + wi) * NCH + c] = 0u;        // left clean for the next batch (no memset per call)
    const int64_t w_lo = r_lo + (int64_t)wi * Lw;          // first row of the window
    if (c == 0) s_n = 0;
    __syncthreads();
    if (a) s_keys[atomicAdd(&s_n, 1)] = ((uint32_t)((int64_t)a - 1 - w_lo) << 9) | (uint32_t)(c + 1);
    if (c == 0) {                                          // end-of-second event inside this window?
        const int64_t ta = t_abs0 + w_lo;
        const int64_t tb = min(t_abs0 + w_lo + Lw, t_abs0 + r_hi);
        const int64_t B = ta <= 0 ? SEC_US : ((ta + SEC_US - 1) / SEC_US) * SEC_US;
        if (B < tb) s_keys[atomicAdd(&s_n, 1)] = ((uint32_t)(B - ta) << 9);          // channel field 0: before all channels
    }
    const uint32_t off = win_off[(size_t)board * (n_win + 1) + wi];
    __syncthreads();
    const int n = s_n;
    if (c == 0) win_off[(size_t)board * (n_win + 1) + wi] = 0u;     // (every thread of the CTA has read it)
    uint64_t *out = words + (size_t)board * words_cap;
    // Phase A, a warp per photon word (two per round): the lanes fetch the baseline rows and the peak window and reduce them
    // to what the word needs -- baseline sum, minimum and its position, the parabola's two neighbours -- in shared memory.
    __shared__ int s_S[NCH + 2], s_vmin[NCH + 2], s_y1[NCH + 2], s_y3[NCH + 2];
    for (int i0 = warp; i0 < n; i0 += 16) {
        uint32_t key[2];
        int vb[2], vw[2], ve[2];
        bool live[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {                      // all loads of both triggers first
            const int i = i0 + 8 * u;
            live[u] = i < n;
            key[u] = live[u] ? s_keys[i] : 0u;
            const int ch = (int)(key[u] & 0x1FFu) - 1;
            vb[u] = 0; vw[u] = 0; ve[u] = 0;
            if (live[u] && ch >= 0) {
                const int16_t *q = ph + (w_lo + (int64_t)(key[u] >> 9)) * NCH + ch;
                if (lane < M) vb[u] = q[-(int64_t)(1 + lane) * NCH];
                if (lane < W) vw[u] = q[(int64_t)lane * NCH];
                if (lane == 0) ve[u] = q[(int64_t)W * NCH];
            }
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            if (!live[u] || (key[u] & 0x1FFu) == 0u) continue;          // (warp-uniform; channel field 0 = the end-of-second word)
            const int S = __reduce_add_sync(0xffffffffu, vb[u]);
            // minimum of the window, first occurrence: (value, index) packed
            unsigned mk = lane < W ? ((unsigned)(vw[u] + 32768) << 5) | (unsigned)lane : 0xFFFFFFFFu;
            mk = __reduce_min_sync(0xffffffffu, mk);
            const int j = (int)(mk & 31u);
            const int w_prev = __shfl_sync(0xffffffffu, vw[u], (j + 31) & 31), w_next = __shfl_sync(0xffffffffu, vw[u], (j + 1) & 31);
            const int b_first = __shfl_sync(0xffffffffu, vb[u], 0), e_last = __shfl_sync(0xffffffffu, ve[u], 0);
            if (lane == 0) {
                const int i = i0 + 8 * u;
                s_S[i] = S; s_vmin[i] = (int)(mk >> 5) - 32768;
                s_y1[i] = j >= 1 ? w_prev : b_first; s_y3[i] = j + 1 < W ? w_next : e_last;
            }
        }
    }
    __syncthreads();
    // Phase B, a thread per photon word: rank among the keys of the window, the parabola in float64, the word.  (One lane
    // of every warp did this for its words one after the other before: the longest serial piece of the kernel.)
    auto finish = [&](int i) {
        const uint32_t k = s_keys[i];
        int rank = 0;
        for (int q = 0; q < n; ++q) rank += s_keys[q] < k;
        const int ch = (int)(k & 0x1FFu) - 1;
        uint64_t word = ~0ull;                              // (channel field 0) the end-of-second word
        if (ch >= 0) {
            const int vmin = s_vmin[i], S = s_S[i];
            const double y1 = (double)s_y1[i], y2 = (double)vmin, y3 = (double)s_y3[i];
            const double den = __dsub_rn(__dadd_rn(y3, y1), __dmul_rn(2.0, y2));
            double y4 = y2;
            if (den != 0.0) {
                const double dy = __dsub_rn(y3, y1);
                y4 = __dsub_rn(y2, __ddiv_rn(__dmul_rn(0.125, __dmul_rn(dy, dy)), den));
            }
            const int peak = (__double2int_rz(__dmul_rn(y4, 0.0625)) + 2048) & 0xFFF;
            const int p1 = (vmin / 16 + 2048) & 0xFFF;
            const int base = (S / (16 * M) + 2048) & 0xFFF;
            const int64_t t = t_abs0 + w_lo + (int64_t)(k >> 9);
            const uint32_t ts = (uint32_t)(t % SEC_US);
            word = ((uint64_t)ch << 56) | ((uint64_t)peak << 44) | ((uint64_t)p1 << 32) | ((uint64_t)base << 20) | ts;
        }
        if ((int64_t)off + rank < words_cap) out[off + rank] = word;
    };
    if (c < n) finish(c);
    if (c == 0 && n > NCH) finish(NCH);                     // 256 triggers + the end-of-second word: one entry more than threads
}


// fold: the rows with even t of odd-bin channels are negated (the hop sign (-1)^(bin (f + 1)) of frame f = t mod Ld);
// *unfoldable is set if a value that would have to be negated is -32768 (the caller then repacks without folding)
__global__ void pack_dds_kernel(const int16_t *I, const int16_t *Q, int n_lut, const float *gain, const int16_t *bins, int fold,
                                uint32_t *out, int *unfoldable) {
    // out[t][m] = lut[(t/2)*512 + 2*((m+154)%256) + (t&1)]   (define_DDS_LUT layout, ROACH_Setup.py:526-530)
    const int Ld = n_lut / 256;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= Ld * NCH) return;
    const int t = idx / NCH, m = idx % NCH;
    const int src = (t >> 1) * 512 + 2 * ((m + 154) & 255) + (t & 1);
    int vi = I[src], vq = Q[src];
    if (fold && (t & 1) == 0 && (bins[m] & 1) && gain[m] != 0.f) {
        if (vi == -32768 || vq == -32768) *unfoldable = 1;
        vi = -vi; vq = -vq;
    }
    out[idx] = gain[m] != 0.f ? ((uint32_t)(uint16_t)(int16_t)vi | ((uint32_t)(uint16_t)(int16_t)vq << 16)) : 0u;
}

// ------------------------------------------------------------------------------------------
// synthetic ADC stream
__device__ __forceinline__ uint64_t mix64(uint64_t x) {
    x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ull; x ^= x >> 27; x *= 0x94d049bb133111ebull; x ^= x >> 31;
    return x;
}
__device__ __forceinline__ float u01(uint64_t h) { return ((h >> 40) + 0.5f) * (1.0f / 16777216.0f); }

// theta[u][i]: phase excursion of tone i at microsecond u (recursive exponential decay of hashed pulses)
__global__ void synth_theta_kernel(float *theta, int n_tones, int64_t n_us, int64_t u_abs0, int board_id,
                                   mkid_synth_params prm) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_tones) return;
    const float decay = expf(-1.0f / prm.tau_us);
    const float p = prm.pulse_rate * 1e-6f;
    const int64_t warm = 512;
    float acc = 0.f;
    for (int64_t u = u_abs0 - warm; u < u_abs0 + n_us; ++u) {
        acc *= decay;
        if (u >= 0) {
            const uint64_t h = mix64(prm.seed ^ mix64(((uint64_t)board_id << 48) ^ ((uint64_t)i << 32) ^ (uint64_t)u));
            if (u01(h) < p) {
                const float depth = prm.deg_lo + (prm.deg_hi - prm.deg_lo) * u01(mix64(h + 1));
                acc -= depth * 0.017453292519943295f;
            }
        }
        if (u >= u_abs0) theta[(u - u_abs0) * n_tones + i] = acc;
    }
}

__global__ void __launch_bounds__(256) synth_adc_kernel(uint32_t *out, int64_t n, int64_t n_abs0, const float *theta,
                                                        const int32_t *tone_bin, const float *tone_amp,
                                                        const float *tone_phase, int board_id, float scale,
                                                        mkid_synth_params prm) {
    extern __shared__ float s_syn[];     // theta row | bins | amp | phase
    float *s_theta = s_syn;
    int *s_bin = reinterpret_cast<int *>(s_syn + prm.n_tones);
    float *s_amp = s_syn + 2 * prm.n_tones;
    float *s_ph = s_syn + 3 * prm.n_tones;
    const int64_t u = blockIdx.x;                  // one microsecond (512 samples) per CTA
    for (int i = threadIdx.x; i < prm.n_tones; i += blockDim.x) {
        s_theta[i] = theta[u * prm.n_tones + i] + tone_phase[i];
        s_bin[i] = tone_bin[i];
        s_amp[i] = tone_amp[i];
    }
    __syncthreads();
    const uint64_t mask = (uint64_t)prm.n_lut - 1;
    const float inv = 6.283185307179586f / (float)prm.n_lut;
#pragma unroll
    for (int s = 0; s < 2; ++s) {
        const int64_t nl = u * 512 + s * 256 + threadIdx.x;
        if (nl >= n) continue;
        const uint64_t na = (uint64_t)(n_abs0 + nl);
        float xr = 0.f, xi = 0.f;
        for (int i = 0; i < prm.n_tones; ++i) {
            const uint32_t frac = (uint32_t)(((uint64_t)(uint32_t)s_bin[i] * na) & mask);
            float sn, cs;
            __sincosf(fmaf((float)frac, inv, s_theta[i]), &sn, &cs);
            xr = fmaf(s_amp[i], cs, xr);
            xi = fmaf(s_amp[i], sn, xi);
        }
        const uint64_t h = mix64(prm.seed ^ mix64(0x9e3779b97f4a7c15ull + ((uint64_t)board_id << 56) + na));
        const float u1 = u01(h), u2 = u01(mix64(h ^ 0xabcdef12345ull));
        const float rad = sqrtf(-2.0f * __logf(u1)) * prm.noise_lsb;
        float sn, cs;
        __sincosf(6.283185307179586f * u2, &sn, &cs);
        int vi = __float2int_rn(fmaf(xr, scale, rad * cs)), vq = __float2int_rn(fmaf(xi, scale, rad * sn));
        vi = max(-2047, min(2047, vi));
        vq = max(-2047, min(2047, vq));
        out[nl] = (uint32_t)(uint16_t)(int16_t)vi | ((uint32_t)(uint16_t)(int16_t)vq << 16);
    }
}

// ------------------------------------------------------------------------------------------
// 12-bit packed ADC stream.  The ROACH ADC delivers 12-bit I and Q; on the host link (PCIe: the bound of the end-to-end
// path) a complex sample travels as 3 bytes instead of the 4 of the int16 pair: little-endian 24-bit group
// I[11:0] | Q[11:0] << 12 (two's complement).  One CTA step = 1024 samples: 768 words staged through shared memory with
// coalesced loads, every thread expands 4 samples (3 words, stride 3: conflict-free) into one 16-byte store.
constexpr int P12_THREADS = 256;
__device__ __forceinline__ uint32_t p12_expand(uint32_t s) {          // 24-bit group -> int16 I | int16 Q << 16
    const uint32_t i16 = (uint32_t)(((int32_t)(s << 20)) >> 20) & 0xFFFFu;
    const uint32_t q16 = (uint32_t)(((int32_t)(s << 8)) >> 20) << 16;
    return i16 | q16;
}
__global__ void __launch_bounds__(P12_THREADS) adc_unpack12_kernel(const uint32_t *__restrict__ packed, uint4 *__restrict__ out,
                                                                    int64_t n_quads) {
    __shared__ uint32_t s_w[3 * P12_THREADS];
    const int tid = threadIdx.x;
    const int64_t n_tiles = (n_quads + P12_THREADS - 1) / P12_THREADS;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t q0 = tile * P12_THREADS;
        const int64_t w0 = 3 * q0, w_end = 3 * n_quads;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const int64_t w = w0 + k * P12_THREADS + tid;
            s_w[k * P12_THREADS + tid] = w < w_end ? packed[w] : 0u;
        }
        __syncthreads();
        const uint32_t a = s_w[3 * tid], b = s_w[3 * tid + 1], c = s_w[3 * tid + 2];
        __syncthreads();
        if (q0 + tid < n_quads)
            out[q0 + tid] = make_uint4(p12_expand(a & 0xFFFFFFu), p12_expand((a >> 24) | ((b & 0xFFFFu) << 8)),
                                       p12_expand((b >> 16) | ((c & 0xFFu) << 16)), p12_expand(c >> 8));
    }
}
// the inverse (tests, input preparation): *bad counts the samples outside [-2048, 2047] (they are stored clipped)
__global__ void __launch_bounds__(P12_THREADS) adc_pack12_kernel(const uint4 *__restrict__ iq, uint32_t *__restrict__ packed,
                                                                  int64_t n_quads, unsigned long long *bad) {
    __shared__ uint32_t s_w[3 * P12_THREADS];
    const int tid = threadIdx.x;
    const int64_t n_tiles = (n_quads + P12_THREADS - 1) / P12_THREADS;
    int n_bad = 0;
    auto group = [&](uint32_t v) -> uint32_t {
        int i = (int16_t)(v & 0xFFFFu), q = (int16_t)(v >> 16);
        const int ic = max(-2048, min(2047, i)), qc = max(-2048, min(2047, q));
        n_bad += (ic != i) + (qc != q);
        return ((uint32_t)ic & 0xFFFu) | (((uint32_t)qc & 0xFFFu) << 12);
    };
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t q0 = tile * P12_THREADS;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (q0 + tid < n_quads) v = iq[q0 + tid];
        const uint32_t g0 = group(v.x), g1 = group(v.y), g2 = group(v.z), g3 = group(v.w);
        s_w[3 * tid] = g0 | (g1 << 24);
        s_w[3 * tid + 1] = (g1 >> 8) | (g2 << 16);
        s_w[3 * tid + 2] = (g2 >> 16) | (g3 << 8);
        __syncthreads();
        const int64_t w0 = 3 * q0, w_end = 3 * n_quads;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const int64_t w = w0 + k * P12_THREADS + tid;
            if (w < w_end) packed[w] = s_w[k * P12_THREADS + tid];
        }
        __syncthreads();
    }
    if (n_bad) atomicAdd(bad, (unsigned long long)n_bad);
}

int ensure(mkid_ctx *ctx, void **p, size_t *cap, size_t bytes) {
    if (*cap >= bytes && *p) return MKID_OK;
    if (*p) { MKID_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); MKID_CUDA(ctx, cudaFree(*p)); *p = nullptr; *cap = 0; }
    const size_t want = bytes + bytes / 8 + 256;
    cudaError_t e = cudaMalloc(p, want);
    if (e != cudaSuccess) return mkid_fail(ctx, MKID_ENOMEM, "device allocation of %zu bytes failed: %s", want, cudaGetErrorString(e));
    *cap = want;
    return MKID_OK;
}

// K5 driver shared by mkid_chan_process and mkid_chan_detect
// MKID_CHAN_TIMING=1: device time of every kernel of the chain on stderr (CUDA events, printed at the next call)
struct ChainTimer {
    static constexpr int N = 8;
    cudaEvent_t ev[N] = {};
    const char *name[N] = {};
    int n = 0;
    bool on = getenv("MKID_CHAN_TIMING") != nullptr;
    void report() {
        if (!on || n < 2) { n = 0; return; }
        cudaEventSynchronize(ev[n - 1]);
        fprintf(stderr, "[mkid chain timing]");
        for (int i = 1; i < n; ++i) { float ms = 0; cudaEventElapsedTime(&ms, ev[i - 1], ev[i]); fprintf(stderr, " %s %.1f us", name[i], ms * 1e3f); }
        fprintf(stderr, "\n");
        n = 0;
    }
    void mark(cudaStream_t st, const char *nm) {
        if (!on || n >= N) return;
        if (!ev[n]) cudaEventCreate(&ev[n]);
        cudaEventRecord(ev[n], st);
        name[n++] = nm;
    }
};
ChainTimer g_timer;

int run_detect(mkid_ctx *ctx, mkid_chan *ch, const int16_t *phase_dev, int64_t rows, int64_t r_lo, int64_t r_hi,
               int64_t t_abs0, uint64_t *words_dev, int64_t words_cap, const uint32_t *mask_ready) {
    const ChanDev &d = ch->d;
    const int B = d.n_boards;
    const int64_t n_groups = (rows + 31) >> 5;
    int rc;
    size_t cap;
    const bool have_mask = mask_ready != nullptr;
    if (!have_mask) {                                   // detection on caller-supplied phase rows: mask set 0 as scratch
        cap = ch->mask_bytes_set[0];
        if ((rc = ensure(ctx, (void **)&ch->mask_set[0], &cap, (size_t)B * n_groups * NCH * 4))) return rc;
        ch->mask_bytes_set[0] = cap;
    }
    const uint32_t *mask = have_mask ? mask_ready : ch->mask_set[0];
    const int n_win = (int)((r_hi - r_lo + d.Lw - 1) / d.Lw);
    cap = ch->acc_bytes;
    if ((rc = ensure(ctx, (void **)&ch->acc, &cap, (size_t)B * n_win * NCH * 4))) return rc;
    if (cap != ch->acc_bytes) ch->detect_clean = false;
    ch->acc_bytes = cap;
    cap = ch->win_bytes;
    if ((rc = ensure(ctx, (void **)&ch->win_cnt, &cap, (size_t)B * (n_win + 1) * 4))) return rc;
    if (cap != ch->win_bytes) ch->detect_clean = false;
    ch->win_bytes = cap;
    if (!ch->detect_clean) {             // first call, reallocation or a call that failed half-way
        MKID_CUDA(ctx, cudaMemsetAsync(ch->acc, 0, ch->acc_bytes, ctx->stream));
        MKID_CUDA(ctx, cudaMemsetAsync(ch->win_cnt, 0, ch->win_bytes, ctx->stream));
    }
    ch->detect_clean = false;
    if (!have_mask) {
        dim3 gc((unsigned)((rows + CAND_ROWS - 1) / CAND_ROWS), B);
        candidates_kernel<<<gc, 256, 0, ctx->stream>>>(phase_dev, rows, d.M, d.thr, ch->mask_set[0]);
        MKID_CHECK_LAUNCH(ctx);
    }
    if (have_mask) {
        for (int q = 0; q < 2; ++q) {
            const int rpc = ch->mask_head_rpc[q];
            if (mask_ready != ch->mask_set[q] || rpc <= 0 || rows <= rpc) continue;
            mask_head_kernel<<<dim3((unsigned)((rows - 1) / rpc), B), NCH, 0, ctx->stream>>>(phase_dev, rows, rpc, d.M, d.thr, ch->mask_set[q]);
            MKID_CHECK_LAUNCH(ctx);
            ch->mask_head_rpc[q] = 0;                     // (done: a second detection of the same rows must not redo it)
            break;
        }
    }
    g_timer.mark(ctx->stream, "memsets");
    // (two-context pipeline: this call runs beside the next channelizer kernel, on the few SMs it leaves)
    if (ch->alternate && have_mask)
        resolve_kernel<128><<<dim3(NCH / RES_CH, B), 64, 0, ctx->stream>>>(mask, rows, r_lo, r_hi, t_abs0, d.L, d.Lw, n_win,
                                                                     d.t_next, ch->acc, ch->win_cnt);
    else
        resolve_kernel<512><<<dim3(NCH / RES_CH, B), 256, 0, ctx->stream>>>(mask, rows, r_lo, r_hi, t_abs0, d.L, d.Lw, n_win,
                                                                      d.t_next, ch->acc, ch->win_cnt);
    MKID_CHECK_LAUNCH(ctx);
    g_timer.mark(ctx->stream, "resolve");
    scan_kernel<<<B, 256, 0, ctx->stream>>>(ch->win_cnt, n_win, r_lo, r_hi, t_abs0, d.Lw, ch->n_words_dev, words_cap, ch->n_words_dev + B);
    MKID_CHECK_LAUNCH(ctx);
    g_timer.mark(ctx->stream, "scan");
    if (d.M >= 1 && d.M <= 32 && d.W >= 1 && d.W <= 32)
        emit_warp_kernel<<<dim3(n_win, B), 256, 0, ctx->stream>>>(phase_dev, rows, ch->acc, ch->win_cnt, n_win, r_lo, r_hi,
                                                                  t_abs0, d.M, d.W, d.Lw, words_dev, words_cap);
    else
        emit_kernel<<<dim3(n_win, B), 256, 0, ctx->stream>>>(phase_dev, rows, ch->acc, ch->win_cnt, n_win, r_lo, r_hi, t_abs0,
                                                             d.M, d.W, d.Lw, words_dev, words_cap);
    MKID_CHECK_LAUNCH(ctx);
    ch->detect_clean = true;             // every entry resolve / scan wrote has been consumed and cleared by emit
    g_timer.mark(ctx->stream, "emit");
    return MKID_OK;
}

}  // namespace

// =============================================================================================
extern "C" int mkid_chan_create(mkid_ctx *ctx, const mkid_chan_params *prm, mkid_chan **out) {
    if (!ctx) return MKID_EINVAL;
    MKID_REQUIRE(ctx, prm && out, "chan_create: NULL argument");
    MKID_REQUIRE(ctx, prm->n_boards >= 1 && prm->n_boards <= 64, "n_boards must be 1..64");
    MKID_REQUIRE(ctx, prm->n_lut >= 512 && (prm->n_lut & (prm->n_lut - 1)) == 0, "n_lut must be a power of two >= 512");
    MKID_REQUIRE(ctx, prm->mean_len >= 4 && prm->mean_len <= 32, "mean_len must be 4..32");
    MKID_REQUIRE(ctx, prm->holdoff >= 32, "holdoff must be >= 32");
    MKID_REQUIRE(ctx, prm->peak_win >= 1 && prm->peak_win <= 60, "peak_win must be 1..60");
    MKID_CUDA(ctx, cudaSetDevice(ctx->device));
    mkid_chan *ch = new mkid_chan();
    ChanDev &d = ch->d;
    const int B = prm->n_boards;
    d.n_boards = B; d.n_lut = prm->n_lut; d.Ld = prm->n_lut / 256;
    d.M = prm->mean_len; d.L = prm->holdoff; d.W = prm->peak_win;
    int lw = 32;
    while (lw * 2 <= d.L && lw * 2 <= 512) lw *= 2;
    d.Lw = lw;
    // history: earliest sample read = 256*(fb_first+1) - 2048 with fb_first >= 2*(-PRE_ROWS) - 24 - 7
    d.H = HOP * (2 * PRE_ROWS + 24 + 8) + WIN;
    auto A = [&](void **p, size_t bytes) -> int {
        cudaError_t e = cudaMalloc(p, bytes);
        return e == cudaSuccess ? 0 : 1;
    };
    int bad = 0;
    bad |= A((void **)&d.window, WIN * 4);
    bad |= A((void **)&d.tw512, 256 * 8);
    bad |= A((void **)&d.tw256, 256 * 8);
    bad |= A((void **)&d.bins, (size_t)B * NCH * 2);
    bad |= A((void **)&d.dds, (size_t)B * d.Ld * NCH * 4);
    bad |= A((void **)&d.gain, (size_t)B * NCH * 4);
    bad |= A((void **)&d.cen_i, (size_t)B * NCH * 4);
    bad |= A((void **)&d.cen_q, (size_t)B * NCH * 4);
    bad |= A((void **)&d.thr, (size_t)B * NCH * 4);
    bad |= A((void **)&d.hist, (size_t)2 * B * (d.H + 2048) * 4);       // two edge buffers: call k reads k & 1 and prepares the other
    bad |= A((void **)&d.t_next, (size_t)B * NCH * 8);
    bad |= A((void **)&ch->n_words_dev, (size_t)(B + 1) * 4);          // [B] word counts + the sticky overflow flag
    if (bad) { mkid_chan_destroy(ctx, ch); return mkid_fail(ctx, MKID_ENOMEM, "chan_create: device allocation failed"); }
    // default window (Hamming-windowed sinc, sum 1, float32) and twiddles, computed in double on the host
    std::vector<float> h(WIN);
    {
        std::vector<double> hd(WIN);
        double sum = 0;
        for (int m = 0; m < WIN; ++m) {
            const double x = (m - (WIN - 1) / 2.0) / NFFT;
            const double sinc = x == 0 ? 1.0 : sin(M_PI * x) / (M_PI * x);
            const double ham = 0.54 - 0.46 * cos(2.0 * M_PI * m / (WIN - 1));
            hd[m] = sinc * ham; sum += hd[m];
        }
        for (int m = 0; m < WIN; ++m) h[m] = (float)(hd[m] / sum);
    }
    std::vector<float2> t512(256), t256(256);
    for (int k = 0; k < 256; ++k) t512[k] = make_float2((float)cos(-2.0 * M_PI * k / 512.0), (float)sin(-2.0 * M_PI * k / 512.0));
    for (int q = 0; q < 16; ++q)
        for (int j = 0; j < 16; ++j) {
            const double a = -2.0 * M_PI * ((j * q) % 256) / 256.0;
            t256[q * 16 + j] = make_float2((float)cos(a), (float)sin(a));
        }
    MKID_CUDA(ctx, cudaMemcpy(d.window, h.data(), WIN * 4, cudaMemcpyHostToDevice));
    MKID_CUDA(ctx, cudaMemcpy(d.tw512, t512.data(), 256 * 8, cudaMemcpyHostToDevice));
    MKID_CUDA(ctx, cudaMemcpy(d.tw256, t256.data(), 256 * 8, cudaMemcpyHostToDevice));
    for (int k = 0; k < FIRT; ++k) d.fir[k] = 0.f;
    *out = ch;
    return mkid_chan_reset(ctx, ch);
}

extern "C" void mkid_chan_destroy(mkid_ctx *ctx, mkid_chan *ch) {
    if (!ch) return;
    if (ctx) { cudaSetDevice(ctx->device); cudaStreamSynchronize(ctx->stream); }
    ChanDev &d = ch->d;
    void *ps[] = {d.window, d.tw512, d.tw256, d.bins, d.dds, d.gain, d.cen_i, d.cen_q, d.thr, d.hist, d.t_next,
                  ch->n_words_dev, ch->halo, ch->phase_set[0], ch->phase_set[1], ch->mask_set[0], ch->mask_set[1], ch->acc, ch->win_cnt, ch->words_dev, ch->in_dev};
    for (void *p : ps) if (p) cudaFree(p);
    for (int i = 0; i < 2 * mkid_chan::EV_RING; ++i) if (ch->ev_k4[i]) cudaEventDestroy(ch->ev_k4[i]);
    delete ch;
}

extern "C" int mkid_chan_set_fir(mkid_ctx *ctx, mkid_chan *ch, const int32_t *fir_int) {
    if (!ctx) return MKID_EINVAL;
    MKID_REQUIRE(ctx, ch && fir_int, "chan_set_fir: NULL argument");
    for (int k = 0; k < FIRT; ++k) {
        MKID_REQUIRE(ctx, fir_int[k] >= -2048 && fir_int[k] <= 2047, "FIR taps are 12-bit two's complement");
        ch->d.fir[k] = (float)((double)fir_int[k] / (2047.0 * 32767.0));
    }
    ch->fir_set = true;
    return MKID_OK;
}

extern "C" int mkid_chan_set_window(mkid_ctx *ctx, mkid_chan *ch, const float *h) {
    if (!ctx) return MKID_EINVAL;
    MKID_REQUIRE(ctx, ch && h, "chan_set_window: NULL argument");
    MKID_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    MKID_CUDA(ctx, cudaMemcpy(ch->d.window, h, WIN * 4, cudaMemcpyDefault));
    return MKID_OK;
}

extern "C" int mkid_chan_set_thresholds(mkid_ctx *ctx, mkid_chan *ch, int32_t board, const int32_t *thresholds) {
    if (!ctx) return MKID_EINVAL;
    MKID_REQUIRE(ctx, ch && thresholds && board >= 0 && board < ch->d.n_boards, "chan_set_thresholds: bad argument");
    MKID_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    MKID_CUDA(ctx, cudaMemcpy(ch->d.thr + (size_t)board * NCH, thresholds, NCH * 4, cudaMemcpyDefault));
    return MKID_OK;
}

extern "C" int mkid_chan_set_board(mkid_ctx *ctx, mkid_chan *ch, int32_t board, const int32_t *bins,
                                   const int16_t *I_dds, const int16_t *Q_dds, const uint8_t *zero_ch,
                                   const int32_t *centers_i, const int32_t *centers_q, const int32_t *thresholds) {
    if (!ctx) return MKID_EINVAL;
    MKID_REQUIRE(ctx, ch && bins && I_dds && Q_dds, "chan_set_board: NULL argument");
    MKID_REQUIRE(ctx, board >= 0 && board < ch->d.n_boards, "chan_set_board: board out of range");
    MKID_CUDA(ctx, cudaSetDevice(ctx->device));
    ChanDev &d = ch->d;
    std::vector<int16_t> b16(NCH);
    std::vector<float> g(NCH), ci(NCH), cq(NCH);
    std::vector<int32_t> th(NCH);
    for (int c = 0; c < NCH; ++c) {
        b16[c] = (int16_t)(((bins[c] % NFFT) + NFFT) % NFFT);
        g[c] = (zero_ch && zero_ch[c]) ? 0.f : 1.f;
        ci[c] = centers_i ? 8.0f * (float)centers_i[c] : 0.f;
        cq[c] = centers_q ? 8.0f * (float)centers_q[c] : 0.f;
        th[c] = thresholds ? thresholds[c] : -25736;
    }
    MKID_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    MKID_CUDA(ctx, cudaMemcpy(d.bins + (size_t)board * NCH, b16.data(), NCH * 2, cudaMemcpyHostToDevice));
    MKID_CUDA(ctx, cudaMemcpy(d.gain + (size_t)board * NCH, g.data(), NCH * 4, cudaMemcpyHostToDevice));
    MKID_CUDA(ctx, cudaMemcpy(d.cen_i + (size_t)board * NCH, ci.data(), NCH * 4, cudaMemcpyHostToDevice));
    MKID_CUDA(ctx, cudaMemcpy(d.cen_q + (size_t)board * NCH, cq.data(), NCH * 4, cudaMemcpyHostToDevice));
    MKID_CUDA(ctx, cudaMemcpy(d.thr + (size_t)board * NCH, th.data(), NCH * 4, cudaMemcpyHostToDevice));
    // DDS LUT: upload in the reference layout, repack on the device to [t][channel] (I | Q<<16)
    int16_t *tmp = nullptr;
    int rc = mkid_scratch(ctx, SCR_AUX1, (size_t)d.n_lut * 4, (void **)&tmp);
    if (rc) return rc;
    MKID_CUDA(ctx, cudaMemcpyAsync(tmp, I_dds, (size_t)d.n_lut * 2, cudaMemcpyDefault, ctx->stream));
    MKID_CUDA(ctx, cudaMemcpyAsync(tmp + d.n_lut, Q_dds, (size_t)d.n_lut * 2, cudaMemcpyDefault, ctx->stream));
    const int total = d.Ld * NCH;
    int *unf = nullptr;
    if ((rc = mkid_scratch(ctx, SCR_AUX2, 16, (void **)&unf))) return rc;
    for (int fold = 1; fold >= 0; --fold) {          // folded hop sign; once more without it if a -32768 stands in the way
        MKID_CUDA(ctx, cudaMemsetAsync(unf, 0, 4, ctx->stream));
        pack_dds_kernel<<<(total + 255) / 256, 256, 0, ctx->stream>>>(tmp, tmp + d.n_lut, d.n_lut, d.gain + (size_t)board * NCH,
                                                                     d.bins + (size_t)board * NCH, fold,
                                                                     d.dds + (size_t)board * d.Ld * NCH, unf);
        MKID_CHECK_LAUNCH(ctx);
        int h_unf = 0;
        MKID_CUDA(ctx, cudaMemcpyAsync(&h_unf, unf, 4, cudaMemcpyDeviceToHost, ctx->stream));
        MKID_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        d.fold[board] = (unsigned char)(fold && !h_unf);
        if (!h_unf) break;
    }
    ch->board_set[board] = true;
    return MKID_OK;
}

extern "C" int mkid_chan_set_f32_phase_out(mkid_ctx *ctx, mkid_chan *ch, float *dev) {
    if (!ctx) return MKID_EINVAL;
    MKID_REQUIRE(ctx, ch, "chan_set_f32_phase_out: NULL");
    MKID_REQUIRE(ctx, dev == nullptr || mkid_is_device_ptr(dev), "f32 phase output must be device memory");
    ch->f32_out = dev;
    return MKID_OK;
}

extern "C" int mkid_chan_reset(mkid_ctx *ctx, mkid_chan *ch) {
    if (!ctx) return MKID_EINVAL;
    MKID_REQUIRE(ctx, ch, "chan_reset: NULL");
    ChanDev &d = ch->d;
    MKID_CUDA(ctx, cudaMemsetAsync(d.hist, 0, (size_t)2 * d.n_boards * (d.H + 2048) * 4, ctx->stream));
    MKID_CUDA(ctx, cudaMemsetAsync(d.t_next, 0, (size_t)d.n_boards * NCH * 8, ctx->stream));
    MKID_CUDA(ctx, cudaMemsetAsync(ch->n_words_dev, 0, (size_t)(d.n_boards + 1) * 4, ctx->stream));
    ch->t_consumed = 0;
    return MKID_OK;
}

namespace {
// resolve / scan / emit on the rows of a finished channelizer call; with n_words (host) the counts and, for a host word
// buffer, the words are fetched (synchronises).  *overflow_need = words per board that would have been needed, or 0.
int detect_and_fetch(mkid_ctx *ctx, mkid_chan *ch, const int16_t *phase_buf, const uint32_t *mask, int64_t rows, int64_t T,
                     int64_t t_abs0, uint64_t *words, int64_t words_cap, int32_t *n_words, int64_t *overflow_need) {
    const int B = ch->d.n_boards;
    int rc;
    size_t cap;
    uint64_t *wdev;
    if (mkid_is_device_ptr(words)) wdev = words;
    else {
        cap = ch->words_bytes;
        if ((rc = ensure(ctx, (void **)&ch->words_dev, &cap, (size_t)B * words_cap * 8))) return rc;
        ch->words_bytes = cap;
        wdev = ch->words_dev;
    }
    if ((rc = run_detect(ctx, ch, phase_buf, rows, RES_LO, RES_LO + T, t_abs0, wdev, words_cap, mask))) return rc;
    if (n_words) {
        MKID_CUDA(ctx, cudaMemcpyAsync(n_words, ch->n_words_dev, (size_t)B * 4, cudaMemcpyDeviceToHost, ctx->stream));
        MKID_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        int64_t mx = 0;
        for (int b = 0; b < B; ++b) mx = std::max<int64_t>(mx, n_words[b]);
        if (wdev != words) {
            const int64_t ncopy = std::min<int64_t>(mx, words_cap);
            for (int b = 0; b < B; ++b)
                if (n_words[b] > 0)
                    MKID_CUDA(ctx, cudaMemcpyAsync(words + (size_t)b * words_cap, wdev + (size_t)b * words_cap,
                                                   (size_t)std::min<int64_t>(n_words[b], ncopy) * 8, cudaMemcpyDeviceToHost,
                                                   ctx->stream));
            MKID_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        }
        *overflow_need = mx > words_cap ? mx : 0;
    }
    return MKID_OK;
}
}  // namespace

namespace {
// one thread that holds its stream for `ns` nanoseconds (two-context pipeline: lets the channelizer kernel of the next batch
// take its SMs before the detection tail of this batch spreads over them)
__global__ void delay_kernel(unsigned ns, const int32_t *n_words_prev, int n_boards, int limit_per_board) {
    // confined to the few SMs K4 leaves, the tail needs about 0.2 ms + 12 us per 1000 photon words: beyond `limit` words per
    // board in the previous batch it would outlast the channelizer kernel, and starting at once (on all SMs, delaying K4 by
    // some 35 us) is the better deal
    long long total = 0;
    for (int b = 0; b < n_boards; ++b) total += n_words_prev[b];
    if (total > (long long)limit_per_board * n_boards) return;
    unsigned long long t0, t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    do {
        __nanosleep(1000);
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    } while (t - t0 < ns);
}
}  // namespace

extern "C" int mkid_chan_set_pipelined(mkid_ctx *ctx, mkid_chan *ch, int32_t on) {
    if (!ctx) return MKID_EINVAL;
    MKID_REQUIRE(ctx, ch, "chan_set_pipelined: NULL");
    ch->alternate = on != 0;
    return MKID_OK;
}

extern "C" int mkid_chan_detect_pending(mkid_ctx *ctx, mkid_chan *ch, uint64_t *words, int64_t words_cap, int32_t *n_words) {
    if (!ctx) return MKID_EINVAL;
    MKID_REQUIRE(ctx, ch && words && words_cap > 0, "chan_detect_pending: NULL argument");
    MKID_REQUIRE(ctx, ch->pending.valid, "chan_detect_pending: no mkid_chan_process(detect = 2) call is waiting for its detection");
    if (!n_words) MKID_REQUIRE(ctx, mkid_is_device_ptr(words), "asynchronous call (n_words == NULL): words must be device memory");
    MKID_CUDA(ctx, cudaSetDevice(ctx->device));
    const mkid_chan::Pending pd = ch->pending;
    ch->pending.valid = false;
    int64_t need = 0;
    if (ch->alternate) {
        // Two-context pipeline.  This call's first kernel (resolve) becomes ready when the channelizer kernel of the batch
        // ends, 13 us before the next channelizer kernel is launched (edge / history copies), and would spread over every SM:
        // the next K4 then waits for SMs (1.24 -> 1.28 ms with 8 boards).  Holding the tail back for 20 us lets K4 take its
        // 144 SMs first and confines the tail to the 4 free ones, where it needs 1.16 ms per batch of 8 boards at 1000
        // triggers per second and channel (resolve in its small-tile shape 0.1 ms, emit 0.9 ms: 1024 per-window CTAs) -- less
        // than K4's 1.24 ms.  Per step, 1 / 2 / 4 / 8 boards: 0.214 -> 0.200, 0.372 -> 0.356,
        // 0.677 -> 0.648, 1.296 -> 1.256 ms.  MKID_TAIL_DELAY_US overrides (0: the tail starts at once).  The delay kernel skips
        // the wait when the previous batch emitted more words than the confined tail can handle in K4's time.
        int delay_us = 20;
        if (const char *e = getenv("MKID_TAIL_DELAY_US")) delay_us = atoi(e);
        // words per board and batch up to which the confined tail fits under K4: K4 takes 4.6 ps per sample, the confined
        // tail 12 ns per word (+ 0.2 ms per batch of 8 boards): about n / 3000 words per board for n samples per board
        int limit = (int)std::min<int64_t>(pd.T * 512 / 3000, 1 << 30);
        if (const char *e = getenv("MKID_TAIL_DELAY_LIMIT")) limit = atoi(e);
        if (delay_us > 0) {
            delay_kernel<<<1, 1, 0, ctx->stream>>>((unsigned)delay_us * 1000u, ch->n_words_dev, ch->d.n_boards, limit);
            MKID_CHECK_LAUNCH(ctx);
        }
    }
    int rc = detect_and_fetch(ctx, ch, ch->phase_set[pd.set], ch->mask_set[pd.set], pd.rows, pd.T, pd.t_abs0, words, words_cap, n_words, &need);
    if (rc) return rc;
    if (need) return mkid_fail(ctx, MKID_EINVAL, "word buffer too small: need %lld per board, have %lld (the words beyond the capacity are lost)",
                               (long long)need, (long long)words_cap);
    return MKID_OK;
}

extern "C" int mkid_chan_process(mkid_ctx *ctx, mkid_chan *ch, const int16_t *iq, int64_t n, int32_t detect,
                                 uint64_t *words, int64_t words_cap, int32_t *n_words, int16_t *phase_out) {
    if (!ctx) return MKID_EINVAL;
    MKID_REQUIRE(ctx, ch && iq, "chan_process: NULL argument");
    ChanDev &d = ch->d;
    MKID_REQUIRE(ctx, n > 0 && n % 2048 == 0, "n must be a positive multiple of 2048 samples (4 output rows)");
    MKID_REQUIRE(ctx, n >= d.H, "n must be at least the history length (59392 samples) per call");
    MKID_REQUIRE(ctx, ch->fir_set, "FIR taps not set (mkid_chan_set_fir)");
    for (int b = 0; b < d.n_boards; ++b) MKID_REQUIRE(ctx, ch->board_set[b], "a board is not configured (mkid_chan_set_board)");
    MKID_REQUIRE(ctx, detect >= 0 && detect <= 2, "detect must be 0, 1 or 2 (2 = mask only, detection by mkid_chan_detect_pending)");
    if (detect == 1) MKID_REQUIRE(ctx, words && words_cap > 0, "detect requested but no word buffer");
    if (detect == 1 && !n_words) MKID_REQUIRE(ctx, mkid_is_device_ptr(words) && mkid_is_device_ptr(iq) && !phase_out,
                                         "asynchronous call (n_words == NULL): iq and words must be device memory, no phase_out");
    MKID_CUDA(ctx, cudaSetDevice(ctx->device));
    const int B = d.n_boards;
    const int64_t T = n / 512, rows = PRE_ROWS + T;
    int rc;
    size_t cap;
    // input
    const uint32_t *in_dev;
    if (mkid_is_device_ptr(iq)) in_dev = (const uint32_t *)iq;
    else {
        cap = ch->in_bytes;
        if ((rc = ensure(ctx, (void **)&ch->in_dev, &cap, (size_t)B * n * 4))) return rc;
        ch->in_bytes = cap;
        MKID_CUDA(ctx, cudaMemcpyAsync(ch->in_dev, iq, (size_t)B * n * 4, cudaMemcpyHostToDevice, ctx->stream));
        in_dev = ch->in_dev;
    }
    const int set = ch->alternate ? (int)(ch->n_calls & 1) : 0;
    cap = ch->phase_rows_set[set] * NCH * 2 * B;
    if (ch->phase_rows_set[set] < (size_t)rows) {
        if ((rc = ensure(ctx, (void **)&ch->phase_set[set], &cap, (size_t)B * rows * NCH * 2))) return rc;
        ch->phase_rows_set[set] = (size_t)rows;
    }
    int16_t *const phase_buf = ch->phase_set[set];
    // K4: equal chunks of rows, one CTA per SM in a single wave
    WsParams w;
    uint32_t *const edge_cur = d.hist + (size_t)((ch->n_calls_edge) & 1) * B * (d.H + 2048);
    uint32_t *const edge_nxt = d.hist + (size_t)((ch->n_calls_edge + 1) & 1) * B * (d.H + 2048);
    ch->n_calls_edge++;
    w.d = d; w.in = in_dev; w.edge = edge_cur; w.edge_next = edge_nxt; w.n = n; w.f0_abs = 2 * ch->t_consumed; w.phase = phase_buf; w.phase_f32 = ch->f32_out; w.rows = rows;
    {
        // two-context pipeline: leave 4 SMs to the detection / decode kernels of the previous batch, which run on the other
        // context's stream at the same time (with 8 boards 144 of 148 SMs are taken either way; with one board per GPU the
        // kernel would otherwise fill every SM and the tail of the previous batch would queue behind it: measured per
        // step, 1 / 2 / 4 boards: 0.276 -> 0.216, 0.432 -> 0.374, 0.747 -> 0.677 ms; 8 or 12 free SMs are no better)
        int reserve = ch->alternate ? 4 : 0;
        if (const char *e = getenv("MKID_K4_RESERVE")) reserve = std::max(0, std::min(atoi(e), ctx->num_sms - 8));   // experiment switch
        int64_t chunks = std::max<int64_t>(1, (int64_t)(ctx->num_sms - reserve) / B);
        while (chunks > 1 && rows / chunks < 128) chunks = (chunks + 1) / 2;
        int64_t rpc = (rows + chunks - 1) / chunks;
        rpc = std::max<int64_t>(32, (rpc + 31) / 32 * 32);
        w.rows_per_chunk = (int)rpc;
        w.chunks_per_board = (int)((rows + rpc - 1) / rpc);
    }
    w.mask = nullptr; w.halo = nullptr; w.head_deferred = 0;
    if (detect) {       // K5c fused into K4: the candidate mask is produced while the phase is in registers
        const int64_t n_groups = (rows + 31) >> 5;
        cap = ch->mask_bytes_set[set];
        if ((rc = ensure(ctx, (void **)&ch->mask_set[set], &cap, (size_t)B * n_groups * NCH * 4))) return rc;
        ch->mask_bytes_set[set] = cap;
        cap = ch->halo_bytes;
        if ((rc = ensure(ctx, (void **)&ch->halo, &cap, (size_t)B * w.chunks_per_board * 32 * NCH * 2))) return rc;
        ch->halo_bytes = cap;
        w.mask = ch->mask_set[set]; w.halo = ch->halo;
        static const bool head_off = getenv("MKID_K4_HEAD") && atoi(getenv("MKID_K4_HEAD")) == 0;      // experiment switch
        w.head_deferred = (!head_off && w.chunks_per_board > 1) ? 1 : 0;
        ch->mask_head_rpc[set] = w.head_deferred ? w.rows_per_chunk : 0;
    }
    g_timer.report();
    g_timer.mark(ctx->stream, "start");
    cudaEvent_t *evp = &ch->ev_k4[2 * (ch->n_calls % mkid_chan::EV_RING)];
    if (!evp[0]) { cudaEventCreate(&evp[0]); cudaEventCreate(&evp[1]); }
    ch->n_calls++;
    {
        const size_t smem = (size_t)(WS_NBUF * 16 * FFT_STRIDE + 256) * sizeof(float2) + (size_t)(WS_ADC_STAGES + WS_DDS_STAGES) * FB * NCH * 4;
        MKID_CUDA(ctx, cudaEventRecord(evp[0], ctx->stream));
        bool all_fold = true;
        for (int b = 0; b < B; ++b) all_fold = all_fold && d.fold[b];
        auto launch = [&](auto kern) -> cudaError_t {
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
            kern<<<dim3(w.chunks_per_board, B), WS_THREADS, smem, ctx->stream>>>(w);
            return cudaSuccess;
        };
        if (ch->f32_out) MKID_CUDA(ctx, all_fold ? launch(channelize_ws_kernel<true, true>) : launch(channelize_ws_kernel<true, false>));
        else MKID_CUDA(ctx, all_fold ? launch(channelize_ws_kernel<false, true>) : launch(channelize_ws_kernel<false, false>));
        MKID_CHECK_LAUNCH(ctx);
    }
    MKID_CUDA(ctx, cudaEventRecord(evp[1], ctx->stream));
    g_timer.mark(ctx->stream, "channelize");
    if (phase_out) {
        for (int b = 0; b < B; ++b)
            MKID_CUDA(ctx, cudaMemcpyAsync(phase_out + (size_t)b * T * NCH, phase_buf + ((size_t)b * rows + PRE_ROWS) * NCH,
                                           (size_t)T * NCH * 2, cudaMemcpyDefault, ctx->stream));
    }
    const int64_t t_abs0 = ch->t_consumed - PRE_ROWS;
    int64_t overflow_need = 0;
    ch->pending = mkid_chan::Pending();
    if (detect == 2) {              // detection deferred to mkid_chan_detect_pending (possibly on another context's stream)
        ch->pending.valid = true; ch->pending.set = set; ch->pending.rows = rows; ch->pending.T = T; ch->pending.t_abs0 = t_abs0;
    } else if (detect) {
        if ((rc = detect_and_fetch(ctx, ch, phase_buf, ch->mask_set[set], rows, T, t_abs0, words, words_cap, n_words, &overflow_need))) return rc;
    }
    ch->t_consumed += T;
    if (!mkid_is_device_ptr(iq) || phase_out) MKID_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    // reported only now: the streaming state (hold-off times, input history, time) is that of a completed call, the
    // words beyond the capacity are lost
    if (overflow_need) return mkid_fail(ctx, MKID_EINVAL, "word buffer too small: need %lld per board, have %lld (the call completed, "
                                        "the words beyond the capacity are lost)", (long long)overflow_need, (long long)words_cap);
    return MKID_OK;
}

extern "C" int mkid_chan_last_kernel_ms(mkid_ctx *ctx, mkid_chan *ch, float *ms) {
    if (!ctx) return MKID_EINVAL;
    MKID_REQUIRE(ctx, ch && ms && ch->n_calls > 0, "chan_last_kernel_ms: no process call yet");
    cudaEvent_t *evp = &ch->ev_k4[2 * ((ch->n_calls - 1) % mkid_chan::EV_RING)];
    MKID_CUDA(ctx, cudaEventSynchronize(evp[1]));
    MKID_CUDA(ctx, cudaEventElapsedTime(ms, evp[0], evp[1]));
    return MKID_OK;
}

extern "C" int mkid_chan_kernel_ms_sum(mkid_ctx *ctx, mkid_chan *ch, int32_t last_n, float *ms_sum) {
    if (!ctx) return MKID_EINVAL;
    MKID_REQUIRE(ctx, ch && ms_sum && last_n >= 1 && last_n <= mkid_chan::EV_RING && last_n <= ch->n_calls,
                 "chan_kernel_ms_sum: last_n must be 1..64 and not exceed the calls made");
    float sum = 0.f;
    for (int64_t k = ch->n_calls - last_n; k < ch->n_calls; ++k) {
        cudaEvent_t *evp = &ch->ev_k4[2 * (k % mkid_chan::EV_RING)];
        float ms = 0.f;
        MKID_CUDA(ctx, cudaEventSynchronize(evp[1]));
        MKID_CUDA(ctx, cudaEventElapsedTime(&ms, evp[0], evp[1]));
        sum += ms;
    }
    *ms_sum = sum;
    return MKID_OK;
}

extern "C" int mkid_chan_overflowed(mkid_ctx *ctx, mkid_chan *ch, int32_t *flag, int32_t clear) {
    if (!ctx) return MKID_EINVAL;
    MKID_REQUIRE(ctx, ch && flag, "chan_overflowed: NULL");
    MKID_CUDA(ctx, cudaMemcpyAsync(flag, ch->n_words_dev + ch->d.n_boards, 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (clear) MKID_CUDA(ctx, cudaMemsetAsync(ch->n_words_dev + ch->d.n_boards, 0, 4, ctx->stream));
    MKID_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return MKID_OK;
}

extern "C" int mkid_chan_n_words_dev(mkid_ctx *ctx, mkid_chan *ch, const int32_t **out) {
    if (!ctx) return MKID_EINVAL;
    MKID_REQUIRE(ctx, ch && out && ch->n_words_dev, "chan_n_words_dev: NULL");
    *out = ch->n_words_dev;
    return MKID_OK;
}

extern "C" int mkid_chan_detect(mkid_ctx *ctx, mkid_chan *ch, const int16_t *phase, int64_t rows, int64_t t_abs0,
                                int64_t *t_next, uint64_t *words, int64_t words_cap, int32_t *n_words) {
    if (!ctx) return MKID_EINVAL;
    MKID_REQUIRE(ctx, ch && phase && words && n_words && words_cap > 0, "chan_detect: NULL argument");
    ChanDev &d = ch->d;
    MKID_REQUIRE(ctx, rows > d.M + d.W + 1, "chan_detect: too few rows");
    MKID_CUDA(ctx, cudaSetDevice(ctx->device));
    const int B = d.n_boards;
    int rc;
    const void *ph_dev;
    if ((rc = mkid_stage_in(ctx, phase, (size_t)B * rows * NCH * 2, SCR_IN, &ph_dev))) return rc;
    if (t_next) MKID_CUDA(ctx, cudaMemcpyAsync(d.t_next, t_next, (size_t)B * NCH * 8, cudaMemcpyDefault, ctx->stream));
    void *w_dev;
    if ((rc = mkid_stage_out(ctx, words, (size_t)B * words_cap * 8, SCR_OUT0, false, &w_dev))) return rc;
    if ((rc = run_detect(ctx, ch, (const int16_t *)ph_dev, rows, d.M, rows - d.W - 1, t_abs0, (uint64_t *)w_dev, words_cap, nullptr))) return rc;
    MKID_CUDA(ctx, cudaMemcpyAsync(n_words, ch->n_words_dev, (size_t)B * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (t_next) MKID_CUDA(ctx, cudaMemcpyAsync(t_next, d.t_next, (size_t)B * NCH * 8, cudaMemcpyDefault, ctx->stream));
    MKID_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if ((rc = mkid_stage_out_finish(ctx, words, (size_t)B * words_cap * 8, w_dev))) return rc;
    for (int b = 0; b < B; ++b)
        if (n_words[b] > words_cap) return mkid_fail(ctx, MKID_EINVAL, "word buffer too small: need %d", n_words[b]);
    return MKID_OK;
}

extern "C" int mkid_synth_adc(mkid_ctx *ctx, const mkid_synth_params *prm, int32_t n_boards, const int32_t *tone_bin,
                              const float *tone_amp, const float *tone_phase, int64_t n, int64_t t_abs0_us, int16_t *out) {
    if (!ctx) return MKID_EINVAL;
    MKID_REQUIRE(ctx, prm && tone_bin && tone_amp && tone_phase && out, "synth_adc: NULL argument");
    MKID_REQUIRE(ctx, prm->n_tones >= 1 && prm->n_tones <= 1024 && n > 0 && n % 512 == 0, "synth_adc: bad sizes");
    MKID_REQUIRE(ctx, prm->n_lut >= 512 && (prm->n_lut & (prm->n_lut - 1)) == 0, "synth_adc: n_lut must be a power of two");
    MKID_CUDA(ctx, cudaSetDevice(ctx->device));
    const int Tn = prm->n_tones;
    const int64_t n_us = n / 512;
    int rc;
    float *theta; int32_t *d_bin; float *d_amp, *d_ph;
    if ((rc = mkid_scratch(ctx, SCR_AUX2, (size_t)n_us * Tn * 4, (void **)&theta))) return rc;
    if ((rc = mkid_scratch(ctx, SCR_AUX3, (size_t)n_boards * Tn * 12, (void **)&d_bin))) return rc;
    d_amp = (float *)(d_bin + (size_t)n_boards * Tn);
    d_ph = d_amp + (size_t)n_boards * Tn;
    MKID_CUDA(ctx, cudaMemcpyAsync(d_bin, tone_bin, (size_t)n_boards * Tn * 4, cudaMemcpyDefault, ctx->stream));
    MKID_CUDA(ctx, cudaMemcpyAsync(d_amp, tone_amp, (size_t)n_boards * Tn * 4, cudaMemcpyDefault, ctx->stream));
    MKID_CUDA(ctx, cudaMemcpyAsync(d_ph, tone_phase, (size_t)n_boards * Tn * 4, cudaMemcpyDefault, ctx->stream));
    void *o_dev;
    if ((rc = mkid_stage_out(ctx, out, (size_t)n_boards * n * 4, SCR_OUT0, false, &o_dev))) return rc;
    // amplitude scale: sum of squares -> 4 sigma headroom
    std::vector<float> amp_h((size_t)n_boards * Tn);
    MKID_CUDA(ctx, cudaMemcpyAsync(amp_h.data(), d_amp, amp_h.size() * 4, cudaMemcpyDeviceToHost, ctx->stream));
    MKID_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (int b = 0; b < n_boards; ++b) {
        double ss = 0;
        for (int i = 0; i < Tn; ++i) ss += (double)amp_h[(size_t)b * Tn + i] * amp_h[(size_t)b * Tn + i];
        const float scale = Tn == 1 ? prm->full_scale / (float)sqrt(ss) : prm->full_scale / (4.0f * (float)sqrt(ss));
        synth_theta_kernel<<<(Tn + 63) / 64, 64, 0, ctx->stream>>>(theta, Tn, n_us, t_abs0_us, b, *prm);
        MKID_CHECK_LAUNCH(ctx);
        synth_adc_kernel<<<(unsigned)n_us, 256, (size_t)Tn * 16, ctx->stream>>>(
            (uint32_t *)o_dev + (size_t)b * n, n, t_abs0_us * 512, theta, d_bin + (size_t)b * Tn, d_amp + (size_t)b * Tn,
            d_ph + (size_t)b * Tn, b, scale, *prm);
        MKID_CHECK_LAUNCH(ctx);
    }
    return mkid_stage_out_finish(ctx, out, (size_t)n_boards * n * 4, o_dev);
}

extern "C" int mkid_adc_unpack12(mkid_ctx *ctx, const void *packed, int64_t n_samples, int16_t *iq) {
    MKID_REQUIRE(ctx, packed && iq && n_samples > 0 && n_samples % 4 == 0, "adc_unpack12: NULL argument or n_samples not a multiple of 4");
    MKID_REQUIRE(ctx, mkid_is_device_ptr(packed) && mkid_is_device_ptr(iq), "adc_unpack12: device pointers required (use mkid_upload_async / mkid_memcpy for the host side)");
    MKID_REQUIRE(ctx, ((uintptr_t)packed & 3) == 0 && ((uintptr_t)iq & 15) == 0, "adc_unpack12: packed must be 4-byte and iq 16-byte aligned");
    const int64_t n_quads = n_samples / 4;
    const int64_t tiles = (n_quads + P12_THREADS - 1) / P12_THREADS;
    const int grid = (int)std::min<int64_t>(tiles, (int64_t)ctx->num_sms * 8);
    adc_unpack12_kernel<<<grid, P12_THREADS, 0, ctx->stream>>>((const uint32_t *)packed, (uint4 *)iq, n_quads);
    MKID_CHECK_LAUNCH(ctx);
    return MKID_OK;
}

extern "C" int mkid_adc_pack12(mkid_ctx *ctx, const int16_t *iq, int64_t n_samples, void *packed, int64_t *n_clipped) {
    MKID_REQUIRE(ctx, packed && iq && n_samples > 0 && n_samples % 4 == 0, "adc_pack12: NULL argument or n_samples not a multiple of 4");
    MKID_REQUIRE(ctx, mkid_is_device_ptr(packed) && mkid_is_device_ptr(iq), "adc_pack12: device pointers required");
    MKID_REQUIRE(ctx, ((uintptr_t)packed & 3) == 0 && ((uintptr_t)iq & 15) == 0, "adc_pack12: packed must be 4-byte and iq 16-byte aligned");
    int rc;
    unsigned long long *bad;
    if ((rc = mkid_scratch(ctx, SCR_AUX5, 8, (void **)&bad))) return rc;
    MKID_CUDA(ctx, cudaMemsetAsync(bad, 0, 8, ctx->stream));
    const int64_t n_quads = n_samples / 4;
    const int64_t tiles = (n_quads + P12_THREADS - 1) / P12_THREADS;
    const int grid = (int)std::min<int64_t>(tiles, (int64_t)ctx->num_sms * 8);
    adc_pack12_kernel<<<grid, P12_THREADS, 0, ctx->stream>>>((const uint4 *)iq, (uint32_t *)packed, n_quads, bad);
    MKID_CHECK_LAUNCH(ctx);
    if (n_clipped) {
        unsigned long long h = 0;
        MKID_CUDA(ctx, cudaMemcpyAsync(&h, bad, 8, cudaMemcpyDeviceToHost, ctx->stream));
        MKID_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        *n_clipped = (int64_t)h;
    }
    return MKID_OK;
}
